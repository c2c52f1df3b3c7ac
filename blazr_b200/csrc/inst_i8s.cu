#define B200Q_FMT FmtI8S
#define B200Q_FAM_ID B200Q_FAM_I8S
#define B200Q_HAS_GGML_REPACK 0
#include "inst_body.cuh"
