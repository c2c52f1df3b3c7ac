#define B200Q_FMT FmtQ4_1
#define B200Q_FAM_ID B200Q_FAM_Q4_1
#define B200Q_HAS_GGML_REPACK 1
#include "inst_body.cuh"
