// Fake <nccl.h> (types only) for the host-emulated build: the library binds NCCL with dlopen and only when world > 1,
// which the emulation never is.
#pragma once
#include <cstddef>
typedef struct EmulNcclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclChar = 0, ncclInt = 2, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclMin = 3 } ncclRedOp_t;
