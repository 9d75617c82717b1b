// resize.cuh — the two cv2.resize calls around the generator call of inference.py, on the GPU (scope row f2):
//   inference.py:126       face = cv2.resize(face, (96, 96))                      crop of a frame -> generator input
//   inference.py:269-271   p = cv2.resize(p.astype(np.uint8), (x2-x1, y2-y1)); f[y1:y2, x1:x2] = p   prediction -> frame
// OpenCV's 8-bit INTER_LINEAR is fixed-point (11-bit coefficients, 32-bit row sums, a 2-stage shift in the vertical
// pass); the arithmetic below is that algorithm, integer for integer, so results are bit-identical to cv2's
// (oracle/pipeline_oracle.py: resize_linear_u8, pinned against cv2 itself).  HBM-bound byte work: one thread per
// destination pixel (3 channels), 4 source pixels read through the read-only path.
#pragma once

#include <stdint.h>

namespace w2l {

struct ResizeAxis { int s0, s1, w0, w1; };

// source index pair and 11-bit weights of destination index d (resize.cpp, resizeGeneric_ coefficient tables)
__device__ __forceinline__ ResizeAxis resize_axis(int d, int dst, int src, bool x_axis) {
    const double inv_scale = (double)dst / (double)src;
    const double scale = 1.0 / inv_scale;
    const double v = __dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    float f = (float)v;
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    ResizeAxis a;
    if (x_axis) {   // columns: out-of-range taps are removed by forcing the fraction to 0
        if (s < 0) { f = 0.0f; s = 0; }
        if (s >= src - 1) { f = 0.0f; s = src - 1; }
        a.s0 = s; a.s1 = min(s + 1, src - 1);
    } else {        // rows: indices are clamped, the weights stay
        a.s0 = min(max(s, 0), src - 1); a.s1 = min(max(s + 1, 0), src - 1);
    }
    a.w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
    a.w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
    return a;
}

__device__ __forceinline__ void resize_pixel(const uint8_t* src, long long row_pitch, const ResizeAxis& ax, const ResizeAxis& ay,
                                             uint8_t* out3) {
    const uint8_t* r0 = src + ay.s0 * row_pitch;
    const uint8_t* r1 = src + ay.s1 * row_pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = (int)__ldg(r0 + ax.s0 * 3 + c) * ax.w0 + (int)__ldg(r0 + ax.s1 * 3 + c) * ax.w1;
        const int h1 = (int)__ldg(r1 + ax.s0 * 3 + c) * ax.w0 + (int)__ldg(r1 + ax.s1 * 3 + c) * ax.w1;
        out3[c] = (uint8_t)((((ay.w0 * (h0 >> 4)) >> 16) + ((ay.w1 * (h1 >> 4)) >> 16) + 2) >> 2);
    }
}

// boxes: [N][5] = (frame index, y1, y2, x1, x2)
__global__ void crop_resize_kernel(const uint8_t* frames, int H, int W, const int* boxes, int N, int S, uint8_t* crops) {
    const long long total = (long long)N * S * S;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % S), oy = (int)((i / S) % S), n = (int)(i / ((long long)S * S));
        const int* b = boxes + 5 * n;
        const int f = b[0], y1 = b[1], y2 = b[2], x1 = b[3], x2 = b[4];
        const ResizeAxis ax = resize_axis(ox, S, x2 - x1, true);
        const ResizeAxis ay = resize_axis(oy, S, y2 - y1, false);
        const uint8_t* src = frames + (((long long)f * H + y1) * W + x1) * 3;
        uint8_t o[3];
        resize_pixel(src, (long long)W * 3, ax, ay, o);
        uint8_t* d = crops + i * 3;
        d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
    }
}

// out[n] = copy of frames[box.f] with the prediction pred[n] (S x S) resized to the box and pasted into it
__global__ void paste_kernel(const uint8_t* pred, int S, const uint8_t* frames, int H, int W, const int* boxes, int N, uint8_t* out) {
    const long long total = (long long)N * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H), n = (int)(i / ((long long)W * H));
        const int* b = boxes + 5 * n;
        const int f = b[0], y1 = b[1], y2 = b[2], x1 = b[3], x2 = b[4];
        uint8_t o[3];
        if (y >= y1 && y < y2 && x >= x1 && x < x2) {
            const ResizeAxis ax = resize_axis(x - x1, x2 - x1, S, true);
            const ResizeAxis ay = resize_axis(y - y1, y2 - y1, S, false);
            resize_pixel(pred + (long long)n * S * S * 3, (long long)S * 3, ax, ay, o);
        } else {
            const uint8_t* s = frames + (((long long)f * H + y) * W + x) * 3;
            o[0] = __ldg(s); o[1] = __ldg(s + 1); o[2] = __ldg(s + 2);
        }
        uint8_t* d = out + i * 3;
        d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
    }
}

}  // namespace w2l
