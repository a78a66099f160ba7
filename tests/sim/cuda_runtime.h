/* cuda_runtime.h -- TEST SCAFFOLDING ONLY (tests/sim): a stand-in for the handful of CUDA runtime calls csrc/vit_api.cu makes, so
 * that the library's HOST code (vit_run's sequential / time-sliced / chunk-pipeline paths, the staging pool, vit_stream_push) can
 * be compiled with a plain C++ compiler and run in the GPU-less authoring container against the kernel SOURCE in the host
 * emulator (tests/emu).  Implemented by tests/sim/sim_runtime.cpp: "device" memory is host memory (poisoned on allocation),
 * streams are in-order queues executed by a small scheduler that runs every kernel as EARLY as its dependencies allow and
 * re-runs a gate-waiting kernel at every gate opening -- an adversarial schedule for the upload protocols.
 * Never part of the product: the product has no CPU path. */
#pragma once
#include <math.h>     /* the real header pulls the math functions in as well */
#include <stddef.h>

typedef enum cudaError { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorPeerAccessAlreadyEnabled = 704 } cudaError_t;
typedef struct SimStream* cudaStream_t;
typedef struct SimEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { enum cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaFuncAttributes { int numRegs; size_t sharedSizeBytes; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2 };

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError(void);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned flags);
cudaError_t cudaMemset(void* p, int v, size_t n);
cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes* a, const void* p);
cudaError_t cudaFuncGetAttributes(struct cudaFuncAttributes* a, const void* f);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, enum cudaMemcpyKind k, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, enum cudaMemcpyKind k, cudaStream_t s);
