// cu_utils.h -- fatal-error helpers with the reference's semantics (message + exit(1);
// /root/reference/src/cu_utils.cpp:4-37).
#ifndef CUPSS_B200_CU_UTILS_H
#define CUPSS_B200_CU_UTILS_H
#include "defines.h"
void check_error(cudaError_t err);
void check_device();
#endif
