/* fake_cabi.c -- TEST SCAFFOLDING ONLY (tests/test_host_harness_cpu.py): a stand-in for the host-buffer part of the C ABI
 * (include/vit_b200.h) that decodes with the golden model (oracle/vit_oracle.h) on the CPU.  LD_PRELOADed in front of
 * libvitb200.so so that the host-side C++ (host/main.cpp, viterbiDF.h, dataflow.h, viterbi.h) can be run end to end in the
 * GPU-less authoring container.  It is never built into, shipped with or loaded by the product: the product has no CPU
 * path.  Only the entry points the ./main host pipeline uses are defined here. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/vit_b200.h"
#include "../../oracle/vit_oracle.h"

struct vit_handle { int options; unsigned long long launches; };
static char g_err[256] = "";

const char* vit_last_error(void) { return g_err; }
void vit_code_parameters(int* k, int* p1, int* p2) {
    unsigned a, b;
    vo_get_polynomials(&a, &b);
    if (k) *k = 7;
    if (p1) *p1 = (int)a;
    if (p2) *p2 = (int)b;
}
int vit_options_valid(int o) {
    int it = o & 0xf, mt = (o >> 4) & 0xf, ot = (o >> 8) & 0xf, cm = (o >> 12) & 0xf;
    return it <= 4 && mt <= 2 && ot <= 1 && cm <= 2 && (o >> 16) == 0 && !(mt == 1 && it == 3) && !(cm == 2 && mt == 2);
}
size_t vit_input_size(int o, size_t n) { return vo_input_size(o & 0xfff, n); }
size_t vit_message_len(int o, size_t n) { return vo_message_len(o & 0xfff, n); }
size_t vit_output_size(int o, size_t n) { return vo_output_size(o & 0xfff, n); }
int vit_create(vit_handle** out, int options, int device, size_t prealloc) {
    (void)device; (void)prealloc;
    if (!vit_options_valid(options)) { snprintf(g_err, sizeof g_err, "unsupported option combination 0x%x", options); return VIT_ERR_OPTIONS; }
    *out = (vit_handle*)calloc(1, sizeof(vit_handle));
    (*out)->options = options;
    return VIT_OK;
}
void vit_destroy(vit_handle* h) { free(h); }
int vit_run(vit_handle* h, const void* in_h, void* out_h, size_t inputNum, float* kernel_ms) {
    int o = h->options;
    if ((o & 0xf000) == 0x1000) o &= 0xfff;          /* -c dpx selects the same core as reg */
    if (vo_decode(o, in_h, out_h, inputNum, 0, 0) != 0) { snprintf(g_err, sizeof g_err, "golden model rejects 0x%x", o); return VIT_ERR_OPTIONS; }
    h->launches++;
    if (kernel_ms) *kernel_ms = 1.0f;                /* not a measurement */
    return VIT_OK;
}
