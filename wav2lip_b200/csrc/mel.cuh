// mel.cuh — fused audio.melspectrogram for sm_100a.
//
// One kernel does what /root/reference/audio.py:45-51 does in six NumPy/librosa passes:
//   pre-emphasis (audio.py:20-23) -> reflect-padded framing + periodic Hann window + 800-point real FFT
//   (audio.py:57-61 -> librosa.stft) -> |.| -> 80-band Slaney mel filterbank (audio.py:92-101)
//   -> 20*log10(max(1e-5, .)) - 20 (audio.py:103-105, :47) -> symmetric normalise + clip (audio.py:110-114).
//
// Numerics follow the reference's dtypes: pre-emphasis, window and FFT in float64 (scipy.lfilter and
// the FFT of a float64 frame), spectrum rounded to complex64, magnitude / mel / log / clip in float32.
// float64 matters: with a loud tone in the frame, a float32 FFT's noise floor (-144 dB re peak) reaches
// the mel bands near the 1e-5 clipping floor and breaks the 1e-4 tolerance; the B200 has the FP64 rate.
//
// Work split: a block owns MEL_FPB consecutive frames (so the (80, F) row-major output is written in
// 16-byte runs), MEL_TPF threads per frame. The real 800-point FFT is a 400-point complex Stockham FFT
// (radix 5,5,4,4 — 800 = 2^5 5^2 is not a power of two and zero-padding would change the result) in
// shared memory, followed by the even/odd split post-pass; the mel product uses the filterbank's
// sparsity (739 non-zeros, <= 27 per band) straight from the magnitudes in shared memory.
#pragma once

#include <stdint.h>

#include "mel_consts.h"

namespace w2l {

constexpr int MEL_FPB = 4;
constexpr int MEL_TPF = 128;
constexpr int MEL_NFFT = 800;
constexpr int MEL_HOP = 200;
constexpr int MEL_BINS = 401;
constexpr int MEL_BANDS = 80;

struct MelParams {
    const float* wav;
    long long L;
    float* mel;          // (80, F) row-major
    long long F;
    const double2* tw;   // [0,401): exp(-2 pi i m / 800) (window + real-FFT post pass); [401, 401+395): per-pass twiddles
                         // T[r][k] = exp(-2 pi i r k / (Ns R)), r-major so that a warp reads consecutive k (no bank conflicts)
    const float* bvals;  // packed non-zero filterbank weights
    const int* boff;     // [80] offset into bvals
    const int* bstart;   // [80] first FFT bin of the band
    const int* blen;     // [80] number of bins
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

template <int R>
__device__ __forceinline__ void butterfly(double2* v);

template <>
__device__ __forceinline__ void butterfly<4>(double2* v) {
    const double2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
    const double2 c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
    v[0] = cadd(a, c);
    v[2] = csub(a, c);
    v[1] = make_double2(b.x + d.y, b.y - d.x);  // b - i d
    v[3] = make_double2(b.x - d.y, b.y + d.x);  // b + i d
}

template <>
__device__ __forceinline__ void butterfly<5>(double2* v) {
    const double c1 = 0.30901699437494742410229341718282;   // cos(2 pi / 5)
    const double c2 = -0.80901699437494742410229341718282;  // cos(4 pi / 5)
    const double s1 = 0.95105651629515357211643933337938;   // sin(2 pi / 5)
    const double s2 = 0.58778525229247312916870595463907;   // sin(4 pi / 5)
    const double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    const double2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const double2 t1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const double2 t2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const double2 u1 = make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const double2 u2 = make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    v[0] = make_double2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = make_double2(t1.x + u1.y, t1.y - u1.x);  // t1 - i u1
    v[4] = make_double2(t1.x - u1.y, t1.y + u1.x);  // t1 + i u1
    v[2] = make_double2(t2.x + u2.y, t2.y - u2.x);
    v[3] = make_double2(t2.x - u2.y, t2.y + u2.x);
}

// Twiddle table layout in shared memory (double2 units)
constexpr int MEL_TW_POST = 0;                 // 401 entries
constexpr int MEL_TW_P2 = 401;                 // radix 5, Ns = 5  : T[r-1][k], r = 1..4, k < 5    (20)
constexpr int MEL_TW_P3 = MEL_TW_P2 + 20;      // radix 4, Ns = 25 : T[r-1][k], r = 1..3, k < 25   (75)
constexpr int MEL_TW_P4 = MEL_TW_P3 + 75;      // radix 4, Ns = 100: T[r-1][k], r = 1..3, k < 100  (300)
constexpr int MEL_TW_TOTAL = MEL_TW_P4 + 300;  // 796

// One Stockham pass of a 400-point FFT: radix R, Ns = product of the radices already applied.
// tw_pass = this pass's r-major twiddle table (nullptr for the first pass, whose twiddles are all 1).
template <int R>
__device__ __forceinline__ void fft400_pass(const double2* in, double2* out, int j, int Ns, const double2* tw_pass) {
    const int k = j % Ns;
    double2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        v[r] = in[j + r * (400 / R)];
        if (r > 0 && tw_pass != nullptr) v[r] = cmul(v[r], tw_pass[(r - 1) * Ns + k]);
    }
    butterfly<R>(v);
    const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) out[j0 + r * Ns] = v[r];
}

__global__ void __launch_bounds__(MEL_FPB* MEL_TPF) mel_kernel(const MelParams p) {
    extern __shared__ uint8_t mel_smem[];
    double2* tw = reinterpret_cast<double2*>(mel_smem);                       // [MEL_TW_TOTAL]
    double2* bufs = tw + MEL_TW_TOTAL;                                        // [FPB][2][400]
    float* mags = reinterpret_cast<float*>(bufs + MEL_FPB * 2 * 400);         // [FPB][404]
    float* outs = mags + MEL_FPB * 404;                                       // [80][FPB]

    const int f = threadIdx.x / MEL_TPF;
    const int tid = threadIdx.x % MEL_TPF;
    const long long t = (long long)blockIdx.x * MEL_FPB + f;
    const bool live = t < p.F;

    for (int i = threadIdx.x; i < MEL_TW_TOTAL; i += blockDim.x) tw[i] = p.tw[i];
    __syncthreads();

    double2* b0 = bufs + f * 800;
    double2* b1 = b0 + 400;
    float* mag = mags + f * 404;

    // ---- frame gather: reflect pad of the PRE-EMPHASISED signal, Hann window, even/odd packing ----
    if (live) {
        for (int i = tid; i < 400; i += MEL_TPF) {
            double s[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int n = 2 * i + e;
                long long j = t * MEL_HOP + n - MEL_NFFT / 2;
                // np.pad(mode="reflect") index folding; clips shorter than n_fft/2 reflect more than once
                while (j < 0 || j >= p.L) j = j < 0 ? -j : 2 * (p.L - 1) - j;
                const double x0 = (double)__ldg(p.wav + j);
                const double y = (j > 0) ? x0 + (-0.97) * (double)__ldg(p.wav + j - 1) : x0;
                const double w = 0.5 - 0.5 * tw[n <= 400 ? n : MEL_NFFT - n].x;  // periodic Hann: 0.5 - 0.5 cos(2 pi n / 800)
                s[e] = w * y;
            }
            b0[i] = make_double2(s[0], s[1]);
        }
    }
    __syncthreads();
    if (live && tid < 80) fft400_pass<5>(b0, b1, tid, 1, nullptr);
    __syncthreads();
    if (live && tid < 80) fft400_pass<5>(b1, b0, tid, 5, tw + MEL_TW_P2);
    __syncthreads();
    if (live && tid < 100) fft400_pass<4>(b0, b1, tid, 25, tw + MEL_TW_P3);
    __syncthreads();
    if (live && tid < 100) fft400_pass<4>(b1, b0, tid, 100, tw + MEL_TW_P4);
    __syncthreads();
    // ---- real-FFT post pass: X[k] = E[k] + W800^k O[k]; round to complex64; magnitude in fp32 ----
    if (live) {
        for (int k = tid; k < MEL_BINS; k += MEL_TPF) {
            const double2 zk = b0[k % 400];
            const double2 zm = b0[(400 - k) % 400];
            const double2 e = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double2 d = make_double2(0.5 * (zk.x - zm.x), 0.5 * (zk.y + zm.y));
            const double2 o = make_double2(d.y, -d.x);  // -i d
            const double2 x = cadd(e, cmul(tw[k], o));
            const float re = (float)x.x, im = (float)x.y;  // complex64 store of librosa.stft
            mag[k] = (float)sqrt((double)re * (double)re + (double)im * (double)im);
        }
    }
    __syncthreads();
    // ---- sparse mel product + dB + normalise/clip, all fp32 as NumPy does on float32 arrays ----
    if (live && tid < MEL_BANDS) {
        const int off = p.boff[tid], st = p.bstart[tid], len = p.blen[tid];
        float s = 0.0f;
        for (int j = 0; j < len; ++j) s = fmaf(__ldg(p.bvals + off + j), mag[st + j], s);
        float db = 20.0f * log10f(fmaxf(1e-5f, s)) - 20.0f;
        float v = 8.0f * ((db + 100.0f) / 100.0f) - 4.0f;
        v = fminf(fmaxf(v, -4.0f), 4.0f);
        outs[tid * MEL_FPB + f] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MEL_BANDS * MEL_FPB; i += blockDim.x) {
        const int m = i / MEL_FPB, ff = i % MEL_FPB;
        const long long tt = (long long)blockIdx.x * MEL_FPB + ff;
        if (tt < p.F) p.mel[(long long)m * p.F + tt] = outs[i];
    }
}

constexpr int kMelSmemBytes = MEL_TW_TOTAL * 16 + MEL_FPB * 800 * 16 + MEL_FPB * 404 * 4 + MEL_BANDS * MEL_FPB * 4;

// ------------------------------------------------------------------------------------------------
// mel_kernel_v2 — the same arithmetic with the 400-point FFT held in REGISTERS (round 2).
//
// The first version round-tripped double2 data through shared memory on every Stockham pass and was shared-memory bound
// (ncu: l1tex 85 %, 0.026 of the HBM roof).  Here 400 = 25 x 16 is split Cooley-Tukey style across threads:
//   step 1  (25 threads per frame, thread = n1): the 16 packed samples z[n1 + 25 n2] are gathered straight from HBM/L1
//           (coalesced across n1), pre-emphasised, windowed (Hann from two table values per thread and compile-time
//           (cos, sin)(n2 pi/8) constants) and transformed by a 16-point FFT (4 x 4) entirely in registers; the result is
//           multiplied by W400^(n1 k2) (a 15-step recurrence from one table value) and written ONCE to shared memory;
//   step 3  (16 threads per frame, thread = k2): 25-point FFT (5 x 5) in registers over n1, then the real-FFT split
//           X[k] = E[k] + W800^k O[k], whose partner Z[400 - k] lives in thread 16 - k2 of the same 16-lane group:
//           warp shuffles, no second shared-memory pass; |X| goes to shared memory as fp32 for the sparse mel product.
// Shared-memory traffic per frame drops from ~70 KB to ~18 KB; all twiddles inside the small FFTs are compile-time
// constants (mel_consts.h).  Same dtypes as before: float64 up to the complex64 rounding of the spectrum, fp32 after.
// ------------------------------------------------------------------------------------------------
constexpr int MEL2_FPB = 5;        // frames per block: 125 of 128 threads busy in step 1, 80 in step 3
constexpr int MEL2_THREADS = 128;
constexpr int kMel2SmemBytes = 404 * 16 + MEL2_FPB * 400 * 16 + MEL2_FPB * 404 * 4 + MEL_BANDS * MEL2_FPB * 4;

__device__ __forceinline__ double2 shfl_d2(double2 v, int src_lane) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src_lane, 16), __shfl_sync(0xffffffffu, v.y, src_lane, 16));
}

__global__ void __launch_bounds__(MEL2_THREADS, 4) mel_kernel_v2(const MelParams p) {
    extern __shared__ uint8_t mel_smem[];
    double2* tw = reinterpret_cast<double2*>(mel_smem);                 // [401] exp(-2 pi i m / 800)  (+3 pad)
    double2* ybuf = tw + 404;                                            // [FPB][16][25]
    float* mags = reinterpret_cast<float*>(ybuf + MEL2_FPB * 400);       // [FPB][404]
    float* outs = mags + MEL2_FPB * 404;                                 // [80][FPB]
    const int tid = threadIdx.x;
    for (int i = tid; i < 401; i += MEL2_THREADS) tw[i] = p.tw[i];
    __syncthreads();

    // ---------------- step 1: thread = (frame f, n1) ----------------
    if (tid < MEL2_FPB * 25) {
        const int f = tid / 25, n1 = tid % 25;
        const long long t = (long long)blockIdx.x * MEL2_FPB + f;
        if (t < p.F) {
            double2 v[16];
            const long long base = t * MEL_HOP - MEL_NFFT / 2 + 2 * n1;
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const long long j0 = base + 50 * n2;
                double y0, y1;
                if (j0 >= 1 && j0 + 1 < p.L) {   // interior: three consecutive samples
                    const double xm = (double)__ldg(p.wav + j0 - 1), x0 = (double)__ldg(p.wav + j0), x1 = (double)__ldg(p.wav + j0 + 1);
                    y0 = x0 + (-0.97) * xm;
                    y1 = x1 + (-0.97) * x0;
                } else {                         // np.pad(mode="reflect") of the PRE-EMPHASISED signal
                    double yy[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        long long j = j0 + e;
                        while (j < 0 || j >= p.L) j = j < 0 ? -j : 2 * (p.L - 1) - j;
                        const double x0 = (double)__ldg(p.wav + j);
                        yy[e] = (j > 0) ? x0 + (-0.97) * (double)__ldg(p.wav + j - 1) : x0;
                    }
                    y0 = yy[0]; y1 = yy[1];
                }
                // periodic Hann 0.5 - 0.5 cos(2 pi n / 800), n = 2 n1 + 50 n2 (+1), from the table (cos is even around 400)
                const int n = 2 * n1 + 50 * n2;
                const double c0 = tw[n <= 400 ? n : MEL_NFFT - n].x;
                const double c1 = tw[n + 1 <= 400 ? n + 1 : MEL_NFFT - n - 1].x;
                v[n2] = make_double2((0.5 - 0.5 * c0) * y0, (0.5 - 0.5 * c1) * y1);
            }
            // 16-point FFT over n2 = 4 a + b  ->  k2 = c + 4 d
            double2 u[4][4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double2 q[4] = {v[b], v[4 + b], v[8 + b], v[12 + b]};
                butterfly<4>(q);
#pragma unroll
                for (int c = 0; c < 4; ++c) u[b][c] = (b * c == 0) ? q[c] : cmul(q[c], kMelW16[b * c]);
            }
            double2 Y[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                double2 q[4] = {u[0][c], u[1][c], u[2][c], u[3][c]};
                butterfly<4>(q);
#pragma unroll
                for (int d = 0; d < 4; ++d) Y[c + 4 * d] = q[d];
            }
            double2* dst = ybuf + f * 400 + n1;
            dst[0] = Y[0];
#pragma unroll
            for (int k2 = 1; k2 < 16; ++k2) {      // W400^(n1 k2) = W800^(2 n1 k2 mod 800), upper half by conjugate symmetry
                const int m = (2 * n1 * k2) % MEL_NFFT;
                double2 w = tw[m <= 400 ? m : MEL_NFFT - m];
                if (m > 400) w.y = -w.y;
                dst[k2 * 25] = cmul(Y[k2], w);
            }
        }
    }
    __syncthreads();

    // ---------------- step 3: thread = (frame f, k2), whole warps so that the shuffles below are convergent ----------------
    if (tid < 96) {
        const int f = tid >> 4, k2 = tid & 15;
        const long long t = (long long)blockIdx.x * MEL2_FPB + f;
        const bool live = f < MEL2_FPB && t < p.F;
        double2 y[25];
#pragma unroll
        for (int n1 = 0; n1 < 25; ++n1) y[n1] = live ? ybuf[f * 400 + k2 * 25 + n1] : make_double2(0.0, 0.0);
        // 25-point FFT over n1 = 5 a + b  ->  k1 = c + 5 d
        double2 u[5][5];
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            double2 q[5] = {y[b], y[5 + b], y[10 + b], y[15 + b], y[20 + b]};
            butterfly<5>(q);
#pragma unroll
            for (int c = 0; c < 5; ++c) u[b][c] = (b * c == 0) ? q[c] : cmul(q[c], kMelW25[b * c]);
        }
        double2 Z[25];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            double2 q[5] = {u[0][c], u[1][c], u[2][c], u[3][c], u[4][c]};
            butterfly<5>(q);
#pragma unroll
            for (int d = 0; d < 5; ++d) Z[c + 5 * d] = q[d];
        }
        // real-FFT split: bin k = k2 + 16 k1 pairs with 400 - k = (16 - k2) + 16 (24 - k1): lane (16 - k2) & 15, register 24 - k1
        const int partner = (16 - k2) & 15;
        const double2 wk2 = live ? tw[k2] : make_double2(1.0, 0.0);
        float* mag = mags + f * 404;
#pragma unroll
        for (int k1 = 0; k1 < 25; ++k1) {
            double2 zp = shfl_d2(Z[24 - k1], partner);
            if (k2 == 0) zp = (k1 == 0) ? Z[0] : Z[25 - k1];    // 400 - 16 k1 = 16 (25 - k1): same lane
            const double2 zk = Z[k1];
            const double2 e = make_double2(0.5 * (zk.x + zp.x), 0.5 * (zk.y - zp.y));
            const double2 dd = make_double2(0.5 * (zk.x - zp.x), 0.5 * (zk.y + zp.y));
            const double2 o = make_double2(dd.y, -dd.x);         // -i d
            const double2 x = cadd(e, cmul(cmul(wk2, kMelW50[k1]), o));
            const float re = (float)x.x, im = (float)x.y;        // complex64 store of librosa.stft
            if (live) mag[k2 + 16 * k1] = (float)sqrt((double)re * (double)re + (double)im * (double)im);
        }
        if (live && k2 == 0) {                                   // Nyquist bin: X[400] = E[0] - O[0] = Re Z[0] - Im Z[0]
            const float re = (float)(Z[0].x - Z[0].y);
            mag[400] = fabsf(re);
        }
    }
    __syncthreads();
    // ---------------- sparse mel product + dB + normalise / clip, fp32 as NumPy does on float32 arrays ----------------
    for (int i = tid; i < MEL_BANDS * MEL2_FPB; i += MEL2_THREADS) {
        const int f = i / MEL_BANDS, m = i % MEL_BANDS;
        const long long t = (long long)blockIdx.x * MEL2_FPB + f;
        if (t >= p.F) continue;
        const float* mag = mags + f * 404;
        const int off = p.boff[m], st = p.bstart[m], len = p.blen[m];
        float s = 0.0f;
        for (int j = 0; j < len; ++j) s = fmaf(__ldg(p.bvals + off + j), mag[st + j], s);
        const float db = 20.0f * log10f(fmaxf(1e-5f, s)) - 20.0f;
        float v = 8.0f * ((db + 100.0f) / 100.0f) - 4.0f;
        v = fminf(fmaxf(v, -4.0f), 4.0f);
        outs[m * MEL2_FPB + f] = v;
    }
    __syncthreads();
    for (int i = tid; i < MEL_BANDS * MEL2_FPB; i += MEL2_THREADS) {
        const int m = i / MEL2_FPB, ff = i % MEL2_FPB;
        const long long tt = (long long)blockIdx.x * MEL2_FPB + ff;
        if (tt < p.F) p.mel[(long long)m * p.F + tt] = outs[i];
    }
}

}  // namespace w2l
