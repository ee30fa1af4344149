// api.cu — the extern "C" surface of libmmbridge.so (include/mmbridge.h): argument checking, error
// text, model lifetime and dispatch to the kernels.  No exceptions cross this boundary.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "mmb_device.cuh"
#include "mmb_internal.h"

namespace mmb {

static thread_local char g_error[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return MMB_OK;
    return fail(MMB_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

// device image of the step table: [n_steps][4] = (bc, cc, sp, 0) then [n_steps][T] temb
static size_t table_floats(int n_steps, int T) { return (size_t)n_steps * (4 + (size_t)T); }

}  // namespace mmb

using namespace mmb;

extern "C" {

int mmb_abi_version(void) { return MMB_ABI_VERSION; }

const char* mmb_last_error(void) { return g_error; }

size_t mmb_epic_packed_floats(const MmbEpicDims* dims) { return dims ? mmb_epic_layout(dims).total : 0; }

int mmb_epic_create(const MmbEpicDims* dims, const float* packed, size_t n_floats, int device, MmbEpicModel** out) {
    if (!dims || !packed || !out) return fail(MMB_EINVAL, "mmb_epic_create: null argument");
    const MmbEpicLayout lo = mmb_epic_layout(dims);
    if (n_floats != lo.total) return fail(MMB_EINVAL, "mmb_epic_create: blob has %zu floats, layout wants %zu", n_floats, lo.total);
    int prev = 0;
    if (int rc = cuda_ok(cudaGetDevice(&prev), "cudaGetDevice")) return rc;
    if (int rc = cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return rc;
    EpicModel* m = new (std::nothrow) EpicModel();
    if (!m) return fail(MMB_ENOMEM, "out of host memory");
    m->dims = *dims;
    m->layout = lo;
    m->device = device;
    m->w = nullptr;
    m->tc_image = nullptr;
    m->tc_image_bytes = 0;
    int rc = cuda_ok(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device), "sm count");
    if (!rc) rc = cuda_ok(cudaMalloc(&m->w, lo.total * sizeof(float)), "cudaMalloc weights");
    if (!rc) rc = cuda_ok(cudaMemcpy(m->w, packed, lo.total * sizeof(float), cudaMemcpyHostToDevice), "weight upload");
    if (!rc && tc_supported(dims, 128)) rc = tc_build_image(m, packed);
    cudaSetDevice(prev);
    if (rc) {
        mmb_epic_destroy(reinterpret_cast<MmbEpicModel*>(m));
        return rc;
    }
    *out = reinterpret_cast<MmbEpicModel*>(m);
    return MMB_OK;
}

void mmb_epic_destroy(MmbEpicModel* handle) {
    EpicModel* m = reinterpret_cast<EpicModel*>(handle);
    if (!m) return;
    if (m->w) cudaFree(m->w);
    if (m->tc_image) cudaFree(m->tc_image);
    delete m;
}

int mmb_epic_forward(const MmbEpicModel* handle, const float* x, const uint8_t* k, const uint8_t* mask,
                     const float* temb, int temb_stride, int B, int N,
                     float* v_out, float* logits_out, float* hidden_out, int precision, void* stream) {
    const EpicModel* m = reinterpret_cast<const EpicModel*>(handle);
    if (!m || !x || !k || !mask || !temb || !v_out || !logits_out) return fail(MMB_EINVAL, "mmb_epic_forward: null argument");
    if (B < 0 || N < 0) return fail(MMB_EINVAL, "mmb_epic_forward: negative size");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (precision == MMB_PREC_FP32)
        return launch_epic_forward_fp32(m, x, k, mask, temb, temb_stride, B, N, v_out, logits_out, hidden_out, s);
    if (precision == MMB_PREC_BF16) {
        if (!m->tc_image || !tc_supported(&m->dims, N))
            return fail(MMB_EUNSUPPORTED, "tcgen05 path is built for H=16, G<=32, Dc+S<=16, head<=16, N<=128; use fp32");
        return launch_epic_forward_tc(m, x, k, mask, temb, temb_stride, B, N, v_out, logits_out, hidden_out, s);
    }
    return fail(MMB_EINVAL, "unknown precision %d", precision);
}

int mmb_bridge_update(float* x, uint8_t* k, uint8_t* mask, const float* v, const float* logits,
                      const float* absorb_logit, const float* u_jump, const float* u_absorb,
                      float dt, float bc, float cc, float sp, int B, int N, int Dc, int S, int flags, void* stream) {
    if (B < 0 || N < 0) return fail(MMB_EINVAL, "mmb_bridge_update: negative size");
    if (Dc < 1 || Dc > 8 || S < 1 || S > 32) return fail(MMB_EINVAL, "mmb_bridge_update: need 1<=Dc<=8, 1<=S<=32");
    if (!mask) return fail(MMB_EINVAL, "mmb_bridge_update: mask is null");
    if (!(flags & MMB_FLAG_NO_EULER) && (!x || !v)) return fail(MMB_EINVAL, "mmb_bridge_update: Euler step needs x and v");
    if (!(flags & MMB_FLAG_NO_JUMP) && (!k || !logits || !u_jump)) return fail(MMB_EINVAL, "mmb_bridge_update: jump needs k, logits, u_jump");
    if ((flags & MMB_FLAG_ABSORBING) && (!absorb_logit || !u_absorb)) return fail(MMB_EINVAL, "mmb_bridge_update: birth needs absorb_logit, u_absorb");
    const size_t P = (size_t)B * (size_t)N;
    if (P == 0) return MMB_OK;
    return launch_bridge_update(x, k, mask, v, logits, absorb_logit, u_jump, u_absorb, StepScalars{dt, bc, cc, sp},
                                P, Dc, S, flags, static_cast<cudaStream_t>(stream));
}

size_t mmb_generate_workspace_bytes(const MmbEpicModel* handle, int B, int N, int precision) {
    const EpicModel* m = reinterpret_cast<const EpicModel*>(handle);
    (void)B; (void)N; (void)precision;
    if (!m) return 0;
    // room for the device image of a step table of up to 4096 steps
    return table_floats(4096, m->dims.dim_time_emb) * sizeof(float);
}

int mmb_generate(const MmbEpicModel* handle, float* x, uint8_t* k, const uint8_t* mask,
                 const MmbStepTable* st, const float* u_jump, uint64_t seed, uint64_t jet_offset,
                 int B, int N, void* workspace, size_t workspace_bytes, int precision, void* stream) {
    const EpicModel* m = reinterpret_cast<const EpicModel*>(handle);
    if (!m || !x || !k || !mask || !st || !workspace) return fail(MMB_EINVAL, "mmb_generate: null argument");
    if (!st->temb || !st->bc || !st->cc) return fail(MMB_EINVAL, "mmb_generate: incomplete step table");
    if (B < 0 || N < 0 || st->n_steps < 0) return fail(MMB_EINVAL, "mmb_generate: negative size");
    const int T = m->dims.dim_time_emb, n = st->n_steps;
    const size_t need = table_floats(n, T) * sizeof(float);
    if (workspace_bytes < need) return fail(MMB_ENOMEM, "mmb_generate: workspace %zu B < %zu B", workspace_bytes, need);
    if (B == 0 || N == 0 || n == 0) return MMB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // stage the table (a few KB); pageable source: the copy is staged before the call returns
    std::vector<float> img(table_floats(n, T));
    for (int i = 0; i < n; ++i) {
        img[(size_t)i * 4 + 0] = st->bc[i];
        img[(size_t)i * 4 + 1] = st->cc[i];
        img[(size_t)i * 4 + 2] = st->sp ? st->sp[i] : 0.0f;
        img[(size_t)i * 4 + 3] = st->t ? st->t[i] : 0.0f;
    }
    memcpy(img.data() + (size_t)n * 4, st->temb, (size_t)n * T * sizeof(float));
    if (int rc = cuda_ok(cudaMemcpyAsync(workspace, img.data(), need, cudaMemcpyHostToDevice, s), "step table upload")) return rc;
    const float* table = static_cast<const float*>(workspace);
    if (precision == MMB_PREC_FP32)
        return launch_generate_fp32(m, x, k, mask, table, n, st->dt, u_jump, seed, jet_offset, B, N, s);
    if (precision == MMB_PREC_BF16) {
        if (!m->tc_image || !tc_supported(&m->dims, N))
            return fail(MMB_EUNSUPPORTED, "tcgen05 path is built for H=16, G<=32, Dc+S<=16, head<=16, N<=128; use fp32");
        return launch_generate_tc(m, x, k, mask, table, n, st->dt, u_jump, seed, jet_offset, B, N, s);
    }
    return fail(MMB_EINVAL, "unknown precision %d", precision);
}

int mmb_philox_uniforms(float* u, uint64_t seed, uint64_t jet_offset, int n_steps, int B, int N, void* stream) {
    if (!u || n_steps < 0 || B < 0 || N < 0) return fail(MMB_EINVAL, "mmb_philox_uniforms: bad argument");
    return launch_philox_uniforms(u, seed, jet_offset, n_steps, B, N, static_cast<cudaStream_t>(stream));
}

int mmb_validation_histograms(const float* x, const uint8_t* k, const uint8_t* mask, int B, int N, int Dc, int S,
                              int bins, float lo, float hi, int max_mult, uint64_t* counts, void* stream) {
    if (!x || !k || !mask || !counts) return fail(MMB_EINVAL, "mmb_validation_histograms: null argument");
    if (B < 0 || N < 0 || Dc < 1 || S < 1 || bins < 1 || max_mult < 0 || !(hi > lo))
        return fail(MMB_EINVAL, "mmb_validation_histograms: bad shape or range");
    if ((size_t)(Dc * bins + S + max_mult + 1) * 4 > 48 * 1024) return fail(MMB_EINVAL, "mmb_validation_histograms: too many bins");
    if (B == 0 || N == 0) return MMB_OK;
    return launch_validation_histograms(x, k, mask, B, N, Dc, S, bins, lo, hi, max_mult,
                                        reinterpret_cast<unsigned long long*>(counts), static_cast<cudaStream_t>(stream));
}

}  // extern "C"
