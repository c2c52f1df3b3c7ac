#define B200Q_FMT FmtIQ4XS
#define B200Q_FAM_ID B200Q_FAM_IQ4_XS
#define B200Q_HAS_GGML_REPACK 1
#include "inst_body.cuh"
