// On-disk depth format of the reference (src/data_utils.py:88-167): a 16-bit PNG holding depth * 256.
//   load_depth                     z = uint16 / 256.0;  z[z <= 0] = 0               (:137-141)
//   load_depth_with_validity_map   ... and v = 1 where z > 0 else z                 (:100-107)
//   save_depth                     uint32(z * 256.0)   (C truncation)               (:165)
// The PNG container is host work (PIL, as in the reference); these two kernels do the element-wise part on the device so the
// host-to-device copy of a sparse depth map is 2 bytes per pixel instead of 4 + 4 (depth and validity).  Bit-exact.
#include "conv_common.cuh"

namespace kbnet {

namespace {

__global__ void depth_u16_decode_kernel(const uint16_t* __restrict__ z16, long long n, float* __restrict__ depth, float* __restrict__ validity) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float z = (float)z16[i] / 256.0f;      // exact: a 16-bit integer times 2^-8
    z = z <= 0.f ? 0.f : z;
    depth[i] = z;
    if (validity) validity[i] = z > 0.f ? 1.f : z;
}

__global__ void depth_u16_encode_kernel(const float* __restrict__ depth, long long n, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = depth[i] * 256.0f;
    // numpy's float32 -> uint32 cast is a C cast: truncation toward zero; out-of-range / NaN inputs are undefined there, clamp here
    out[i] = v >= 4294967040.f ? 0xFFFFFFFFu : (v > 0.f ? (uint32_t)v : 0u);
}

}  // namespace

int depth_u16_decode_launch(const uint16_t* z16, long long n, float* depth, float* validity, cudaStream_t stream) {
    if (n == 0) return 0;
    depth_u16_decode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(z16, n, depth, validity);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

int depth_u16_encode_launch(const float* depth, long long n, uint32_t* out, cudaStream_t stream) {
    if (n == 0) return 0;
    depth_u16_encode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(depth, n, out);
    KB_CUDA(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace kbnet
