/* TEST INFRASTRUCTURE ONLY -- the handful of NCCL types csrc/gfmd_b200.cu names (it dlopens
 * the library itself); the emulated build has one rank and never calls them. */
#pragma once
#include <cstddef>
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
