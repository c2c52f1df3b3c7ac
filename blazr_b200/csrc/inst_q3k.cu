#define B200Q_FMT FmtQ3K
#define B200Q_FAM_ID B200Q_FAM_Q3_K
#define B200Q_HAS_GGML_REPACK 1
#include "inst_body.cuh"
