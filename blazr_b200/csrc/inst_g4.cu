#define B200Q_FMT FmtG4
#define B200Q_FAM_ID B200Q_FAM_G4
#define B200Q_HAS_GGML_REPACK 0
#include "inst_body.cuh"
