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

}  // namespace w2l
