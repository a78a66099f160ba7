// sim_stubs.cpp -- TEST SCAFFOLDING ONLY: the rest of the C ABI (include/vit_b200.h) for tests/sim/libvitsim.so, so that the
// Python mirror of the host class (which binds every entry point when it loads the library) can be run over the host-path
// simulation.  The device-memory helpers work (on the stand-in runtime's "device" memory); everything that needs a real GPU
// kernel of csrc/vit_synth.cu or the multi-GPU code of csrc/vit_mg.cu reports that it does not exist here.
#include <cuda_runtime.h>
#include <cstring>

#include "../../include/vit_b200.h"

int vit_set_error(int code, const char* msg);     // csrc/vit_api.cu

#define NOT_HERE(name) return vit_set_error(VIT_ERR_CUDA, name ": not available in the host-path simulation")

extern "C" {
int vit_dev_alloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
void vit_dev_free(void* p) { cudaFree(p); }
int vit_dev_sync(void) { cudaStreamSynchronize(nullptr); return VIT_OK; }
int vit_dev_set(int d) { return d == 0 ? VIT_OK : VIT_ERR_CUDA; }
int vit_dev_count(void) { return 1; }
int vit_dev_copy_to_host(void* dst, const void* src, size_t n) { cudaStreamSynchronize(nullptr); memcpy(dst, src, n); return VIT_OK; }
int vit_dev_copy_from_host(void* dst, const void* src, size_t n) { cudaStreamSynchronize(nullptr); memcpy(dst, src, n); return VIT_OK; }
int vit_host_alloc(void** p, size_t n) { return cudaHostAlloc(p, n ? n : 1, 0) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
void vit_host_free(void* p) { cudaFreeHost(p); }

int vit_synth_device(int, size_t, unsigned, int, double, int, void*, void*, void*) { NOT_HERE("vit_synth_device"); }
int vit_synth_device_ex(int, size_t, unsigned, int, double, int, int, void*, void*, void*) { NOT_HERE("vit_synth_device_ex"); }
int vit_count_errors_synth_device(int, const void*, size_t, unsigned, int, unsigned long long*, void*) { NOT_HERE("vit_count_errors_synth_device"); }
int vit_count_errors_device(int, const void*, const void*, size_t, unsigned long long*, void*) { NOT_HERE("vit_count_errors_device"); }
int vit_depuncture_device(int, const void*, size_t, unsigned, unsigned, unsigned, void*, size_t, void*) { NOT_HERE("vit_depuncture_device"); }

int vit_comm_available(void) { return 0; }
int vit_comm_nccl_version(void) { return 0; }
int vit_comm_get_unique_id(void*) { NOT_HERE("vit_comm_get_unique_id"); }
int vit_comm_init_rank(vit_comm**, int, int, const void*, int) { NOT_HERE("vit_comm_init_rank"); }
int vit_comm_init_all(vit_comm**, int, const int*) { NOT_HERE("vit_comm_init_all"); }
void vit_comm_destroy(vit_comm*) {}
int vit_comm_rank(const vit_comm*) { return 0; }
int vit_comm_size(const vit_comm*) { return 1; }
int vit_comm_barrier(vit_comm*) { return VIT_OK; }
int vit_comm_stream_wait(vit_comm*, void*) { return VIT_OK; }
int vit_comm_mark(vit_comm*, int) { return VIT_OK; }
int vit_comm_stream_wait_mark(vit_comm*, int, void*) { return VIT_OK; }
void* vit_comm_stream(vit_comm*) { return nullptr; }
void vit_shard_range(size_t, int, int, size_t* first, size_t* count) { if (first) *first = 0; if (count) *count = 0; }
int vit_shard_owner(size_t, int, size_t) { return 0; }
int vit_comm_shared_alloc(vit_comm*, void**, size_t, int) { NOT_HERE("vit_comm_shared_alloc"); }
int vit_comm_gatherv(vit_comm*, int, const void*, void*, const size_t*, const size_t*, int, void*) { NOT_HERE("vit_comm_gatherv"); }
int vit_job_create(vit_job**, vit_comm*, int, const vit_job_config*) { NOT_HERE("vit_job_create"); }
int vit_job_run(vit_job*, vit_job_result*) { NOT_HERE("vit_job_run"); }
const void* vit_job_gathered(const vit_job*, size_t*) { return nullptr; }
int vit_job_stream_range(const vit_job*, size_t*, size_t*) { NOT_HERE("vit_job_stream_range"); }
int vit_job_stream_errors(const vit_job*, unsigned long long*, size_t) { NOT_HERE("vit_job_stream_errors"); }
void vit_job_destroy(vit_job*) {}
}
