#define B200Q_FMT FmtQ2K
#define B200Q_FAM_ID B200Q_FAM_Q2_K
#define B200Q_HAS_GGML_REPACK 1
#include "inst_body.cuh"
