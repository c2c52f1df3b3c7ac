#define B200Q_FMT FmtTQ2_0
#define B200Q_FAM_ID B200Q_FAM_TQ2_0
#define B200Q_HAS_GGML_REPACK 1
#include "inst_body.cuh"
