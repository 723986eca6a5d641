// On-device action codec: the step immediately after the policy heads in the rollout loop and immediately before
// them in the BC data path.  Pure table look-ups / integer logic, one thread per action:
//
//   to_env   : joint policy action (buttons index 0..8640, camera index 0..120) -> the 20 MineRL button flags + the 2 camera angles
//              (lib/action_mapping.py:215-225 `to_factored`, camera-meta nulling :219-222; lib/actions.py:154-169 `policy2env`,
//              mu-law un-discretisation :96-102 as an 11-entry float64 table built by the host with the reference's own formula).
//   from_env : 20 button flags + 2 camera angles -> (buttons index, camera index, is-null flag)
//              (lib/actions.py:171-178 `env2policy` with the mu-law quantiser :82-94 as float64 bin thresholds found by bisection on
//              the host formula, hence bit-identical binning; lib/action_mapping.py:193-213 `from_factored` incl. the mutually
//              exclusive groups :65-99 and the inventory override).
#pragma once
#include "common.cuh"

namespace vpt {

constexpr int kNumButtons = 20;

// out: [n][22] 8-byte words: 20 x int64 button flags, 2 x float64 camera angles (bit pattern) -> ONE device-to-host copy per step
__global__ void __launch_bounds__(256) codec_to_env_kernel(const long long* __restrict__ buttons, const long long* __restrict__ camera,
                                                            const uint8_t* __restrict__ lut_btn, const uint8_t* __restrict__ lut_cam_off,
                                                            const double* __restrict__ cam_lut, int nbins, int njoint, long long n,
                                                            long long* __restrict__ out, int* __restrict__ bad) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long b = buttons[i], c = camera[i];
    if (b < 0 || b >= njoint || c < 0 || c >= (long long)nbins * nbins) {
        atomicAdd(bad, 1);
        b = 0;
        c = 0;
    }
    long long* o = out + i * (kNumButtons + 2);
#pragma unroll
    for (int k = 0; k < kNumButtons; ++k) o[k] = (long long)lut_btn[b * kNumButtons + k];
    int cy = (int)(c / nbins), cx = (int)(c % nbins);
    if (lut_cam_off[b]) cy = cx = nbins / 2;  // camera meta action off -> null camera (lib/action_mapping.py:219-222)
    o[kNumButtons] = __double_as_longlong(cam_lut[cy]);
    o[kNumButtons + 1] = __double_as_longlong(cam_lut[cx]);
}

__device__ __forceinline__ int camera_bin(double v, const double* __restrict__ thr, int nbins) {
    int k = 0;  // number of thresholds <= v  (thr ascending, nbins - 1 of them); NaN compares false -> bin 0
    for (int j = 0; j < nbins - 1; ++j) k += (v >= thr[j]) ? 1 : 0;
    return k;
}

// btn: [n][20] int64 flags; cam: [n][2] float64; strides: [9] group strides of the joint index; out: [n][3] int64 (buttons, camera, is_null)
__global__ void __launch_bounds__(256) codec_from_env_kernel(const long long* __restrict__ btn, const double* __restrict__ cam, const double* __restrict__ thr,
                                                              int nbins, const long long* __restrict__ strides, long long inventory_idx, long long n,
                                                              long long* __restrict__ out) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long* b = btn + i * kNumButtons;
    // BUTTONS order (lib/actions.py:21-33): 0 attack 1 back 2 forward 3 jump 4 left 5 right 6 sneak 7 sprint 8 use 9 drop 10 inventory 11.. hotbar.1-9
    auto on = [&](int k) { return b[k] != 0; };
    long long hot = 0;
    for (int k = 0; k < 9; ++k)
        if (on(11 + k)) hot = k + 1;                 // the later button wins (lib/action_mapping.py:85-99)
    auto pair = [&](int first, int second, bool cancel) -> long long {
        const bool a = on(first), c2 = on(second);
        if (cancel && a && c2) return 0;              // forward+back / left+right together mean neither
        return c2 ? 2 : (a ? 1 : 0);
    };
    const long long fb = pair(2, 1, true), lr = pair(4, 5, true), ss = pair(7, 6, false);
    const int null_bin = nbins / 2;
    const int cy = camera_bin(cam[2 * i], thr, nbins), cx = camera_bin(cam[2 * i + 1], thr, nbins);
    const bool cam_null = (cy == null_bin) && (cx == null_bin);
    long long joint = hot * strides[0] + fb * strides[1] + lr * strides[2] + ss * strides[3] + (on(8) ? 1 : 0) * strides[4] +
                      (on(9) ? 1 : 0) * strides[5] + (on(0) ? 1 : 0) * strides[6] + (on(3) ? 1 : 0) * strides[7] + (cam_null ? 0 : 1) * strides[8];
    long long cidx = (long long)cy * nbins + cx;
    if (b[10] == 1) {  // inventory overrides everything (lib/action_mapping.py:205-209)
        joint = inventory_idx;
        cidx = (long long)null_bin * nbins + null_bin;
    }
    bool any = false;
    for (int k = 0; k < kNumButtons; ++k) any |= on(k);
    out[3 * i] = joint;
    out[3 * i + 1] = cidx;
    out[3 * i + 2] = (!any && cam_null) ? 1 : 0;     // agent.py:176-180 null-action filter
}

}  // namespace vpt

extern "C" int vpt_codec_to_env(const int64_t* buttons, const int64_t* camera, const uint8_t* lut_btn, const uint8_t* lut_cam_off, const double* cam_lut,
                                int32_t nbins, int32_t njoint, int64_t n, int64_t* out, int32_t* bad, void* stream) {
    using namespace vpt;
    VPT_CHECK(buttons && camera && lut_btn && lut_cam_off && cam_lut && out && bad && n > 0 && nbins > 0, "vpt_codec_to_env: bad argument");
    launch_k(codec_to_env_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const long long*>(buttons), reinterpret_cast<const long long*>(camera), lut_btn, lut_cam_off, cam_lut, nbins, njoint, n,
        reinterpret_cast<long long*>(out), bad);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int vpt_codec_from_env(const int64_t* buttons, const double* camera, const double* thresholds, int32_t nbins, const int64_t* strides,
                                  int64_t inventory_idx, int64_t n, int64_t* out, void* stream) {
    using namespace vpt;
    VPT_CHECK(buttons && camera && thresholds && strides && out && n > 0 && nbins > 1, "vpt_codec_from_env: bad argument");
    launch_k(codec_from_env_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const long long*>(buttons), camera, thresholds, nbins, reinterpret_cast<const long long*>(strides), inventory_idx, n,
        reinterpret_cast<long long*>(out));
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
