// api.cu — introspection entry points and the host-buffer set-abstraction call of libpn2_b200.
#include "pn2_common.cuh"

namespace pn2 {
unsigned long long g_launch_count = 0;

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct SaLayout {
    size_t xyz, new_xyz, fps_idx, idx, cnt, grouped, dev_ws, dev_ws_bytes, total;
};

static SaLayout sa_layout(int b, int n, int m, int nsample) {
    SaLayout L;
    size_t off = 0;
    L.xyz = off;     off = align_up(off + sizeof(float) * (size_t)b * n * 3, 256);
    L.new_xyz = off; off = align_up(off + sizeof(float) * (size_t)b * m * 3, 256);
    L.fps_idx = off; off = align_up(off + sizeof(int) * (size_t)b * m, 256);  // also the channel between the two overlapped kernels
    L.idx = off;     off = align_up(off + sizeof(int) * (size_t)b * m * nsample, 256);
    L.cnt = off;     off = align_up(off + sizeof(int) * (size_t)b * m, 256);
    L.grouped = off; off = align_up(off + sizeof(float) * (size_t)b * m * nsample * 3, 256);
    L.dev_ws_bytes = pn2_sa_layer_device_workspace_bytes(b, n, m, nsample);  // 0 on the overlapped path
    L.dev_ws = off;  off = align_up(off + L.dev_ws_bytes, 256);
    L.total = off;
    return L;
}
}  // namespace pn2

extern "C" {

int pn2_api_version(void) { return PN2_API_VERSION; }

const char* pn2_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

unsigned long long pn2_launch_count(void) { return __atomic_load_n(&pn2::g_launch_count, __ATOMIC_RELAXED); }

size_t pn2_sa_layer_workspace_bytes(int b, int n, int m, int nsample) {
    if (b <= 0 || n <= 0 || m <= 0 || nsample <= 0) return 0;
    return pn2::sa_layout(b, n, m, nsample).total;
}

int pn2_sa_layer_host(int b, int n, int m, float radius, int nsample, const float* h_xyz, float* h_new_xyz,
                      int* h_idx, int* h_pts_cnt, float* h_grouped_xyz, void* workspace, size_t workspace_bytes,
                      void* stream) {
    using namespace pn2;
    if (b <= 0 || n <= 0 || m <= 0 || nsample <= 0 || !(radius > 0.f)) return (int)cudaErrorInvalidValue;
    if (!h_xyz || !workspace) return (int)cudaErrorInvalidValue;
    const SaLayout L = sa_layout(b, n, m, nsample);
    if (workspace_bytes < L.total) return (int)cudaErrorInvalidValue;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return (int)cudaErrorMisalignedAddress;
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    float* d_xyz = reinterpret_cast<float*>(ws + L.xyz);
    float* d_new = reinterpret_cast<float*>(ws + L.new_xyz);
    int* d_fps = reinterpret_cast<int*>(ws + L.fps_idx);
    int* d_idx = reinterpret_cast<int*>(ws + L.idx);
    int* d_cnt = reinterpret_cast<int*>(ws + L.cnt);
    float* d_grp = h_grouped_xyz ? reinterpret_cast<float*>(ws + L.grouped) : nullptr;  // not wanted: not computed

    cudaError_t e = cudaMemcpyAsync(d_xyz, h_xyz, sizeof(float) * (size_t)b * n * 3, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    int rc = pn2_sa_layer_device(b, n, m, radius, nsample, d_xyz, d_fps, d_new, d_idx, d_cnt, d_grp, /*center=*/0,
                                 L.dev_ws_bytes ? ws + L.dev_ws : nullptr, L.dev_ws_bytes, stream);
    if (rc) return rc;
    if (h_new_xyz) {
        e = cudaMemcpyAsync(h_new_xyz, d_new, sizeof(float) * (size_t)b * m * 3, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (h_idx) {
        e = cudaMemcpyAsync(h_idx, d_idx, sizeof(int) * (size_t)b * m * nsample, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (h_pts_cnt) {
        e = cudaMemcpyAsync(h_pts_cnt, d_cnt, sizeof(int) * (size_t)b * m, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return (int)e;
    }
    if (h_grouped_xyz) {
        e = cudaMemcpyAsync(h_grouped_xyz, d_grp, sizeof(float) * (size_t)b * m * nsample * 3, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // extern "C"
