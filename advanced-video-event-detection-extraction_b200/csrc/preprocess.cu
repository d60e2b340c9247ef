// K1: uint8 HWC frames -> normalised bf16 patch rows (or fp32 CHW), bit-exact with the reference chain
//   MemoryManager.resize_frame_for_memory (cv2.resize INTER_AREA to fit 512x512; src/utils/memory_manager.py:299-322)
//   -> open_clip image_transform: PIL Resize(224, BICUBIC, antialias) -> CenterCrop(224) -> ToTensor -> Normalize
//      (applied at src/models/openclip_model.py:165-174,188-193)
// restated in oracle/preprocess_ref.py.  Three small kernels, each computing only what the crop needs:
//   A  area shrink      (fp32 taps, x pass then y pass, rint)            frame  -> mid1 uint8
//   B  horizontal pass  (Pillow 22-bit fixed point, uint8 intermediate)  mid1   -> mid2 uint8 [rows, S, 3]
//   C  vertical pass + ToTensor/Normalize lookup + patch-major bf16 (or fp32 CHW) store
// The coefficient tables are built on the host with the same double-precision arithmetic as
// OpenCV's computeResizeAreaTab and Pillow's precompute_coeffs, cached per frame geometry.
#include <math.h>

#include <vector>

#include "internal.h"
#include "ptx.cuh"

using namespace b200;

namespace {

struct AxisTaps {          // generic "dst <- sum of src taps" table for one axis
    std::vector<int> start;   // [n_out] first source index
    std::vector<int> cnt;     // [n_out]
    std::vector<float> wf;    // [n_out * stride] float weights (area)
    std::vector<int> wi;      // [n_out * stride] fixed-point weights (Pillow)
    int stride = 0;
};

// OpenCV computeResizeAreaTab (modules/imgproc/src/resize.cpp)
AxisTaps area_taps(int ssize, int dsize) {
    AxisTaps t;
    const double inv_scale = dsize / static_cast<double>(ssize);
    const double scale = 1.0 / inv_scale;
    t.stride = static_cast<int>(ceil(scale)) + 2;
    t.start.assign(dsize, 0);
    t.cnt.assign(dsize, 0);
    t.wf.assign(static_cast<size_t>(dsize) * t.stride, 0.f);
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale;
        const double fsx2 = fsx1 + scale;
        const double cell = fmin(scale, ssize - fsx1);
        int sx1 = static_cast<int>(ceil(fsx1)), sx2 = static_cast<int>(floor(fsx2));
        sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
        sx1 = sx1 < sx2 ? sx1 : sx2;
        int n = 0;
        float* w = &t.wf[static_cast<size_t>(dx) * t.stride];
        int first = -1;
        if (sx1 - fsx1 > 1e-3) {
            first = sx1 - 1;
            w[n++] = static_cast<float>((sx1 - fsx1) / cell);
        }
        for (int sx = sx1; sx < sx2; ++sx) {
            if (first < 0) first = sx;
            w[n++] = static_cast<float>(1.0 / cell);
        }
        if (fsx2 - sx2 > 1e-3) {
            if (first < 0) first = sx2;
            w[n++] = static_cast<float>(fmin(fmin(fsx2 - sx2, 1.0), cell) / cell);
        }
        t.start[dx] = first < 0 ? 0 : first;
        t.cnt[dx] = n;
    }
    return t;
}

double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
double triangle_filter(double x) {
    if (x < 0.0) x = -x;
    if (x < 1.0) return 1.0 - x;
    return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc (src/libImaging/Resample.c), full-image box.
AxisTaps pillow_taps(int in_size, int out_size, bool bicubic) {
    AxisTaps t;
    const double fsupport = bicubic ? 2.0 : 1.0;
    const double scale = static_cast<double>(in_size) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = fsupport * filterscale;
    const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
    t.stride = ksize;
    t.start.assign(out_size, 0);
    t.cnt.assign(out_size, 0);
    t.wi.assign(static_cast<size_t>(out_size) * ksize, 0);
    std::vector<double> k(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = static_cast<int>(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double arg = (x + xmin - center + 0.5) * ss;
            const double w = bicubic ? bicubic_filter(arg) : triangle_filter(arg);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
        }
        int* wi = &t.wi[static_cast<size_t>(xx) * ksize];
        for (int x = 0; x < xmax; ++x) {
            if (k[x] < 0)
                wi[x] = static_cast<int>(-0.5 + k[x] * (1 << 22));
            else
                wi[x] = static_cast<int>(0.5 + k[x] * (1 << 22));
        }
        t.start[xx] = xmin;
        t.cnt[xx] = xmax;
    }
    return t;
}

// Device-side view of one axis table, restricted to outputs [o0, o0+n).
struct DevTaps {
    const int* start;
    const int* cnt;
    const float* wf;
    const int* wi;
    int stride;
};

struct Plan {
    int H = 0, W = 0, mode = 0;
    bool has_a = false, has_b = false, has_c = false;
    bool a_fast = false;     // integer scale factors in both axes (OpenCV ResizeAreaFast path)
    bool a_seq = false;      // vertical area taps visit the source rows in sequence (strip walker usable)
    bool a_int = false;      // integer-exact area arithmetic proven equal to the fp32 evaluation (see AhIntParams)
    int a_dx = 0, a_dy = 0, a_div_shift = 0;
    uint32_t a_div_mul = 0;
    int a_max_cx = 0;        // largest horizontal tap count of the area stage
    int b_max_cnt = 0;       // largest tap count of the Pillow horizontal pass
    int c_max_cnt = 0;       // largest tap count of the Pillow vertical pass (rows [top, top+S))
    int a_fx = 1, a_fy = 1;
    int h1 = 0, w1 = 0;      // size after the area shrink
    int nw = 0, nh = 0;      // size after the Pillow resize
    int left = 0, top = 0;   // centre crop
    int ry0 = 0, ry1 = 0;    // rows of the stage-B input/outputs that stage C needs
    int rx0 = 0, rx1 = 0;    // columns of the stage-A output that stage B needs
    int sx0 = 0, sx1 = 0, sy0 = 0, sy1 = 0;   // window of the SOURCE frame that the first stage reads
    DevTaps ax{}, ay{}, bx{}, cy{};
    const uint32_t* aq = nullptr;   // [w1][8] packed IDP.2A weights of the vertical-first area kernel (a_int only)
    size_t mid1_per_frame = 0, mid2_per_frame = 0;
    std::vector<void*> dev;  // device allocations holding the tables
    AxisTaps h_ay;           // host copy of the vertical area taps (strip row programs are built from it)
    AxisTaps h_ax, h_bx;     // host copies of the horizontal area / Pillow taps (fragment tables of the IMMA kernel)
    struct MmaTables {       // device tables of area_hpass_mma_kernel for one alignment of the source window
        const void *a1 = nullptr, *kb1 = nullptr, *b2 = nullptr, *gmeta = nullptr, *ap = nullptr, *kbp = nullptr;
        int ntx = 0, npt = 0, ngroups = 0, nb = 0, seg = 0, pitch1 = 0, pitchA = 0;
        bool ok = false;
    };
    mutable std::map<int, MmaTables> mma;       // keyed by delta = (window start address) & 15
    struct StripTable { const void* rinfo = nullptr; const void* meta = nullptr; int nstrips = 0, max_rows = 0; };
    mutable std::map<int, StripTable> strips;   // rows per strip -> row programs of the fused area kernels (device)
};

template <class T>
const T* upload(b200clip_handle* h, Plan& p, const std::vector<T>& v, int& rc) {
    if (v.empty()) return nullptr;
    void* d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(T)) != cudaSuccess) { rc = B200CLIP_E_NOMEM; return nullptr; }
    if (cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
        rc = B200CLIP_E_CUDA;
        return nullptr;
    }
    p.dev.push_back(d);
    h->allocs.push_back(d);
    return static_cast<const T*>(d);
}

DevTaps upload_taps(b200clip_handle* h, Plan& p, const AxisTaps& t, int& rc) {
    DevTaps d{};
    d.start = upload(h, p, t.start, rc);
    d.cnt = upload(h, p, t.cnt, rc);
    d.wf = upload(h, p, t.wf, rc);
    d.wi = upload(h, p, t.wi, rc);
    d.stride = t.stride;
    return d;
}

void taps_range(const AxisTaps& t, int o0, int o1, int& lo, int& hi) {
    lo = 1 << 30; hi = 0;
    for (int o = o0; o < o1; ++o) {
        lo = t.start[o] < lo ? t.start[o] : lo;
        const int e = t.start[o] + t.cnt[o];
        hi = e > hi ? e : hi;
    }
}

std::map<uint64_t, Plan>& plans(b200clip_handle* h) {
    if (!h->pre_plans) h->pre_plans = new std::map<uint64_t, Plan>();
    return *static_cast<std::map<uint64_t, Plan>*>(h->pre_plans);
}

int py_round_half_even(double v) {  // Python round()
    return static_cast<int>(nearbyint(v));
}

}  // namespace

// ---------------------------------------------------------------------------------------------- kernels
// All three stages use a 2-D grid: blockIdx.y = frame, blockIdx.x * blockDim.x + threadIdx.x = work item inside the
// frame (32-bit index math only -- 64-bit div/mod per thread used to dominate the instruction count).

// Aligned-word fetch of NB contiguous bytes starting at an arbitrary address: NW+1 aligned 32-bit loads, realigned
// with funnel shifts into u[0..NW).  Words past the last needed byte are not dereferenced (clamped).
template <int NW>
__device__ __forceinline__ void load_bytes_aligned(const uint8_t* p, int nbytes, uint32_t (&u)[NW]) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    const uint32_t sh = static_cast<uint32_t>(addr & 3) * 8;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~uintptr_t(3));
    const int last = static_cast<int>(((addr & 3) + static_cast<uintptr_t>(nbytes) - 1) >> 2);
    uint32_t w[NW + 1];
#pragma unroll
    for (int k = 0; k <= NW; ++k) w[k] = __ldg(wp + (k < last ? k : last));
#pragma unroll
    for (int k = 0; k < NW; ++k) u[k] = __funnelshift_r(w[k], w[k + 1], sh);
}
// byte n (compile-time) of a realigned word array as float, exactly: splice into the mantissa of 2^23, subtract 2^23
template <int NW>
__device__ __forceinline__ float byte_as_float(const uint32_t (&u)[NW], int n) {
    return __fadd_rn(__uint_as_float(__byte_perm(u[n >> 2], 0x4B000000u, 0x7650u | (n & 3))), -8388608.0f);
}
template <int NW>
__device__ __forceinline__ int byte_as_int(const uint32_t (&u)[NW], int n) {
    return static_cast<int>(__byte_perm(u[n >> 2], 0u, 0x4440u | (n & 3)));
}

// A: area shrink, generic tap counts.  One thread per output pixel (3 channels); taps accumulated in OpenCV's order
// with unfused fp32 multiply/add.
__global__ void __launch_bounds__(256)
area_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, uint8_t* __restrict__ dst,
            int64_t dst_frame_stride, int oy0, int ny, int ox0, int nx, DevTaps ax, DevTaps ay) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(ny * nx)) return;
    const int y = id / nx, x = id - y * nx;
    const int64_t f = blockIdx.y;
    const int dx = ox0 + x, dy = oy0 + y;
    const int sx0 = ax.start[dx], cx = ax.cnt[dx];
    const int sy0 = ay.start[dy], cy = ay.cnt[dy];
    const float* wx = ax.wf + dx * ax.stride;
    const float* wy = ay.wf + dy * ay.stride;
    const uint8_t* base = src + f * frame_stride + sx0 * 3;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int j = 0; j < cy; ++j) {
        const uint8_t* row = base + static_cast<int64_t>(sy0 + j) * row_stride;
        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
        for (int i = 0; i < cx; ++i) {
            const float a = wx[i];
            b0 = __fadd_rn(b0, __fmul_rn(static_cast<float>(row[i * 3 + 0]), a));
            b1 = __fadd_rn(b1, __fmul_rn(static_cast<float>(row[i * 3 + 1]), a));
            b2 = __fadd_rn(b2, __fmul_rn(static_cast<float>(row[i * 3 + 2]), a));
        }
        const float beta = wy[j];
        if (j == 0) {
            s0 = __fmul_rn(beta, b0); s1 = __fmul_rn(beta, b1); s2 = __fmul_rn(beta, b2);
        } else {
            s0 = __fadd_rn(s0, __fmul_rn(beta, b0));
            s1 = __fadd_rn(s1, __fmul_rn(beta, b1));
            s2 = __fadd_rn(s2, __fmul_rn(beta, b2));
        }
    }
    uint8_t* o = dst + f * dst_frame_stride + (y * nx + x) * 3;
    o[0] = static_cast<uint8_t>(min(max(__float2int_rn(s0), 0), 255));
    o[1] = static_cast<uint8_t>(min(max(__float2int_rn(s1), 0), 255));
    o[2] = static_cast<uint8_t>(min(max(__float2int_rn(s2), 0), 255));
}

// A, fast path for <= 5 horizontal taps (scale <= 4, e.g. 1080p/720p -> 512 wide): per source row the <= 15 source
// bytes come from aligned words, the tap weights live in registers (zero padded: adding +0.0f leaves the fp32
// accumulator bit-identical).  Accumulation order and rounding are exactly those of area_kernel.
__global__ void __launch_bounds__(256)
area_kernel_w5(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, uint8_t* __restrict__ dst,
               int64_t dst_frame_stride, int oy0, int ny, int ox0, int nx, DevTaps ax, DevTaps ay) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(ny * nx)) return;
    const int y = id / nx, x = id - y * nx;
    const int64_t f = blockIdx.y;
    const int dx = ox0 + x, dy = oy0 + y;
    const int sx0 = ax.start[dx], cx = ax.cnt[dx];
    const int sy0 = ay.start[dy], cy = ay.cnt[dy];
    const float* wxp = ax.wf + dx * ax.stride;
    const float* wy = ay.wf + dy * ay.stride;
    float wx[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) wx[i] = i < cx ? wxp[i] : 0.f;
    // Packed fp32x2 arithmetic (FADD2/FMUL2: two independent round-to-nearest ops per instruction, bit-identical to
    // the scalar sequence): channels 0/1 of a tap travel as one pair, channel 2 of taps (0,1) and (2,3) as pairs.
    const f32x2_t kMagic = f2_pack(-8388608.0f, -8388608.0f);
    f32x2_t w01[5], w2[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) w01[i] = f2_pack(wx[i], wx[i]);
    w2[0] = f2_pack(wx[0], wx[1]);
    w2[1] = f2_pack(wx[2], wx[3]);
    const uint8_t* base = src + f * frame_stride + static_cast<int64_t>(sy0) * row_stride + sx0 * 3;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    auto magic = [](const uint32_t (&u)[4], int n) {   // byte n spliced into the mantissa of 2^23
        return __uint_as_float(__byte_perm(u[n >> 2], 0x4B000000u, 0x7650u | (n & 3)));
    };
    for (int j = 0; j < cy; ++j) {
        uint32_t u[4];
        load_bytes_aligned<4>(base + j * row_stride, cx * 3, u);
        // conversions and products run packed (two per instruction); the ACCUMULATION stays scalar: ptxas contracts a
        // packed mul.rn + add.rn pair into FFMA2 (single rounding) even with -fmad=false, which would break parity
        float pr[15];
#pragma unroll
        for (int i = 0; i < 5; ++i)
            f2_unpack(f2_mul(f2_add(f2_pack(magic(u, i * 3), magic(u, i * 3 + 1)), kMagic), w01[i]), pr[i * 3], pr[i * 3 + 1]);
        f2_unpack(f2_mul(f2_add(f2_pack(magic(u, 2), magic(u, 5)), kMagic), w2[0]), pr[2], pr[5]);
        f2_unpack(f2_mul(f2_add(f2_pack(magic(u, 8), magic(u, 11)), kMagic), w2[1]), pr[8], pr[11]);
        pr[14] = __fmul_rn(__fadd_rn(magic(u, 14), -8388608.0f), wx[4]);
        float b0 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            b0 = __fadd_rn(b0, pr[i * 3]);
            b1 = __fadd_rn(b1, pr[i * 3 + 1]);
            b2 = __fadd_rn(b2, pr[i * 3 + 2]);
        }
        const float beta = wy[j];
        if (j == 0) {
            s0 = __fmul_rn(beta, b0); s1 = __fmul_rn(beta, b1); s2 = __fmul_rn(beta, b2);
        } else {
            s0 = __fadd_rn(s0, __fmul_rn(beta, b0));
            s1 = __fadd_rn(s1, __fmul_rn(beta, b1));
            s2 = __fadd_rn(s2, __fmul_rn(beta, b2));
        }
    }
    uint8_t* o = dst + f * dst_frame_stride + (y * nx + x) * 3;
    o[0] = static_cast<uint8_t>(min(max(__float2int_rn(s0), 0), 255));
    o[1] = static_cast<uint8_t>(min(max(__float2int_rn(s1), 0), 255));
    o[2] = static_cast<uint8_t>(min(max(__float2int_rn(s2), 0), 255));
}

// A, strip walker for <= 5 horizontal taps and word-aligned rows (the 1080p/720p -> 512 case of the bench).  One
// thread owns one output column over a strip of output rows and walks DOWN the source rows once, exactly like
// OpenCV's ResizeArea_Invoker: each source row is reduced horizontally once (buf) and feeds the output row(s) it
// overlaps, so a boundary row shared by two output rows is converted and multiplied once instead of twice.
// Arithmetic (bit-identical to the scalar mul.rn/add.rn sequence of area_kernel):
//   * byte -> fp32 -> times alpha in ONE fma: the byte is spliced into the mantissa of 2^23 (m = 2^23 + b, exact) and
//     fma(m, alpha, -(2^23 * alpha)) rounds the exact value b*alpha once, which is mul.rn(float(b), alpha);
//   * products that must not be contracted with the following add are written fma(x, y, negzero) with negzero = -0.0f
//     passed as a kernel ARGUMENT: ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (single rounding) even with
//     -fmad=false, but it cannot fold an fma whose addend it does not know;  x*y + (-0) == x*y in round-to-nearest;
//   * channels 0/1 travel as one packed f32x2 pair (FFMA2/FADD2), channel 2 is scalar.
// The next source row is prefetched while the current one is reduced.
template <int DEPTH>   // source rows in flight ahead of the one being reduced
__global__ void __launch_bounds__(128)
area_strip_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int row_words, uint8_t* __restrict__ dst,
                  int64_t dst_frame_stride, int oy0, int ny, int ox0, int nx, int rows_per_strip, int nstrips, int row_max,
                  DevTaps ax, DevTaps ay, float negzero) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(nstrips * nx)) return;
    const int strip = id / nx, x = id - strip * nx;
    const int64_t f = blockIdx.y;
    const int dx = ox0 + x;
    const int sx0 = ax.start[dx], cx = ax.cnt[dx];
    const float* wxp = ax.wf + dx * ax.stride;
    float wx[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) wx[i] = i < cx ? __ldg(wxp + i) : 0.f;
    // weight pairs and their -(2^23 * w) addends: channels (0,1) of tap i; channel 2 of taps (0,1), (2,3), 4
    f32x2_t w01[5], c01[5], w2[2], c2[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        w01[i] = f2_pack(wx[i], wx[i]);
        c01[i] = f2_pack(-8388608.0f * wx[i], -8388608.0f * wx[i]);
    }
    w2[0] = f2_pack(wx[0], wx[1]); c2[0] = f2_pack(-8388608.0f * wx[0], -8388608.0f * wx[1]);
    w2[1] = f2_pack(wx[2], wx[3]); c2[1] = f2_pack(-8388608.0f * wx[2], -8388608.0f * wx[3]);
    const float w24 = wx[4], c24 = -8388608.0f * wx[4];
    const f32x2_t nz2 = f2_pack(negzero, negzero);

    const uintptr_t addr = reinterpret_cast<uintptr_t>(src + f * frame_stride + sx0 * 3);
    const uint32_t sh = static_cast<uint32_t>(addr & 3) * 8;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~uintptr_t(3));
    const int last = static_cast<int>(((addr & 3) + static_cast<uintptr_t>(cx * 3) - 1) >> 2);   // last word needed
    const int o1 = min(1, last), o2 = min(2, last), o3 = min(3, last), o4 = min(4, last);
    auto load_row = [&](int row, uint32_t (&w)[5]) {
        const uint32_t* p = wp + static_cast<int64_t>(row) * row_words;
        w[0] = __ldg(p); w[1] = __ldg(p + o1); w[2] = __ldg(p + o2); w[3] = __ldg(p + o3); w[4] = __ldg(p + o4);
    };
    auto magic = [](const uint32_t (&u)[4], int n) {   // 2^23 + byte n
        return __uint_as_float(__byte_perm(u[n >> 2], 0x4B000000u, 0x7650u | (n & 3)));
    };

    const int dy_a = oy0 + strip * rows_per_strip;
    const int dy_b = min(dy_a + rows_per_strip, oy0 + ny);
    uint32_t nxt[DEPTH][5];
    int have = __ldg(ay.start + dy_a) - 1;          // row held in (b01, b2); rows are visited in sequence
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load_row(min(have + 1 + d, row_max), nxt[d]);
    f32x2_t b01 = 0;
    float b2 = 0.f;
    uint8_t* o = dst + f * dst_frame_stride + (static_cast<int64_t>(dy_a - oy0) * nx + x) * 3;
    for (int dy = dy_a; dy < dy_b; ++dy, o += nx * 3) {
        const int sy0 = __ldg(ay.start + dy), cy = __ldg(ay.cnt + dy);
        const float* wy = ay.wf + dy * ay.stride;
        f32x2_t s01 = 0;     // 0 + beta*buf == beta*buf exactly (all terms are >= +0)
        float s2 = 0.f;
        for (int j = 0; j < cy; ++j) {
            if (sy0 + j != have) {
                have = sy0 + j;
                uint32_t u[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) u[k] = __funnelshift_r(nxt[0][k], nxt[0][k + 1], sh);
#pragma unroll
                for (int d = 0; d + 1 < DEPTH; ++d)
#pragma unroll
                    for (int k = 0; k < 5; ++k) nxt[d][k] = nxt[d + 1][k];
                load_row(min(have + DEPTH, row_max), nxt[DEPTH - 1]);
                b01 = f2_fma(f2_pack(magic(u, 0), magic(u, 1)), w01[0], c01[0]);
#pragma unroll
                for (int i = 1; i < 5; ++i)
                    b01 = f2_add(b01, f2_fma(f2_pack(magic(u, i * 3), magic(u, i * 3 + 1)), w01[i], c01[i]));
                float p0, p1, p2, p3;
                f2_unpack(f2_fma(f2_pack(magic(u, 2), magic(u, 5)), w2[0], c2[0]), p0, p1);
                f2_unpack(f2_fma(f2_pack(magic(u, 8), magic(u, 11)), w2[1], c2[1]), p2, p3);
                const float p4 = __fmaf_rn(magic(u, 14), w24, c24);
                b2 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3), p4);
            }
            const float beta = __ldg(wy + j);
            s01 = f2_add(s01, f2_fma(f2_pack(beta, beta), b01, nz2));
            s2 = __fadd_rn(s2, __fmaf_rn(beta, b2, negzero));
        }
        float s0, s1;
        f2_unpack(s01, s0, s1);
        o[0] = static_cast<uint8_t>(min(max(__float2int_rn(s0), 0), 255));
        o[1] = static_cast<uint8_t>(min(max(__float2int_rn(s1), 0), 255));
        o[2] = static_cast<uint8_t>(min(max(__float2int_rn(s2), 0), 255));
    }
}

// A (integer scale factors): OpenCV ResizeAreaFast -- integer box sum times float(1/area), rint;
// the 2x2 uint8 case is (a+b+c+d+2)>>2.
__global__ void __launch_bounds__(256)
area_fast_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, uint8_t* __restrict__ dst,
                 int64_t dst_frame_stride, int oy0, int ny, int ox0, int nx, int fx, int fy) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(ny * nx)) return;
    const int y = id / nx, x = id - y * nx;
    const int64_t f = blockIdx.y;
    const float inv_area = 1.0f / static_cast<float>(fx * fy);
    const uint8_t* base = src + f * frame_stride + static_cast<int64_t>(oy0 + y) * fy * row_stride +
                          static_cast<int64_t>(ox0 + x) * fx * 3;
    int s0 = 0, s1 = 0, s2 = 0;
    for (int j = 0; j < fy; ++j) {
        const uint8_t* row = base + j * row_stride;
        for (int i = 0; i < fx; ++i) { s0 += row[i * 3]; s1 += row[i * 3 + 1]; s2 += row[i * 3 + 2]; }
    }
    uint8_t* o = dst + f * dst_frame_stride + (y * nx + x) * 3;
    if (fx == 2 && fy == 2) {
        o[0] = static_cast<uint8_t>((s0 + 2) >> 2);
        o[1] = static_cast<uint8_t>((s1 + 2) >> 2);
        o[2] = static_cast<uint8_t>((s2 + 2) >> 2);
    } else {
        o[0] = static_cast<uint8_t>(min(max(__float2int_rn(__fmul_rn(static_cast<float>(s0), inv_area)), 0), 255));
        o[1] = static_cast<uint8_t>(min(max(__float2int_rn(__fmul_rn(static_cast<float>(s1), inv_area)), 0), 255));
        o[2] = static_cast<uint8_t>(min(max(__float2int_rn(__fmul_rn(static_cast<float>(s2), inv_area)), 0), 255));
    }
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= 22;
    return static_cast<uint8_t>(min(max(v, 0), 255));
}

// A+B fused, bulk-copy fed (the 1080p/720p bench path).  One CTA per (strip of area-output rows, frame):
//   * a producer warp streams the strip's source rows -- only the [xb0, xb0+seg) byte window the crop needs -- into a
//     shared-memory ring with 1-D bulk async copies (cp.async.bulk, mbarrier complete_tx), NSTAGE rows in flight per
//     CTA without holding registers;
//   * consumer thread x walks down the rows for area-output column ox0+x exactly like area_strip_kernel (same
//     arithmetic, bit-identical), reading its <= 15 bytes per row from the ring;
//   * each finished area row (uint8) is parked in shared memory and consumer threads t < S immediately apply Pillow's
//     horizontal fixed-point pass to it and store row y of mid2 -- the area image never travels through HBM.
constexpr int AH_NSTAGE = 8;
struct AhRowInfo {      // how one source row of the strip feeds the vertical accumulation
    float beta_cur;     // weight into the output row being accumulated
    float beta_next;    // weight into the NEXT output row when the row straddles two (else 0)
    int finish;         // 1: this row is the last tap of the output row being accumulated
    int pad;
};
// INTX: integer-exact area arithmetic.  When every area weight is a multiple of 1/Dx (1/Dy) -- 1080p and 720p -> 512
// wide give Dx = Dy = 15 and 5 -- the exact value of an output sample is N / (Dx*Dy) with N an integer; for odd Dx*Dy
// it is never closer than 1/(2*Dx*Dy) to a rounding tie, while OpenCV's fp32 evaluation is within
// (taps_x + taps_y + 6) * 255 * 2^-24 of it.  The host enables INTX only when that bound (with a 2x margin) is below
// the tie distance, so rint(fp32 result) == floor((2N + D) / 2D) for every input: the horizontal taps then run as
// byte dot products (IDP.4A, 4 bytes per instruction, no byte->float conversions).  Verified bit-exact against cv2
// in tests/test_gpu_preprocess.py like the fp32 path.
// PX: area-output columns per consumer thread (tid, tid + ncons): the per-row pipeline bookkeeping (barrier wait,
// release, address updates) is paid once for PX columns.  The horizontal Pillow pass likewise owns up to PX output
// columns per thread.
struct AhIntParams { int dx, dy, d, div_shift; uint32_t div_mul; };
template <int MAXT, int MINB, bool INTX, int PX>
__global__ void __launch_bounds__(MAXT, MINB)
area_hpass_bulk_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride,
                       uint8_t* __restrict__ mid2, int64_t mid2_frame_stride, int oy0, int ny, int ox0, int nx,
                       int rows_per_strip, int max_rows, int xb0, int seg_bytes, int stage_bytes, int arow_pitch,
                       int left, int S, DevTaps ax, DevTaps ay, DevTaps bx, float negzero, AhIntParams ip) {
    static_assert(PX == 1 || INTX, "two columns per thread only with the integer-exact arithmetic (registers)");
    extern __shared__ __align__(128) uint8_t ah_smem[];
    const int ncons = blockDim.x - 32;                 // consumer threads; the last warp is the producer
    const int tid = threadIdx.x;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ah_smem);
    uint64_t* empty_bar = full_bar + AH_NSTAGE;
    uint8_t* ring = ah_smem + 2 * AH_NSTAGE * sizeof(uint64_t);
    uint8_t* arow = ring + AH_NSTAGE * stage_bytes;
    AhRowInfo* rinfo = reinterpret_cast<AhRowInfo*>(arow + 2 * arow_pitch);

    const int strip = blockIdx.x;
    const int64_t f = blockIdx.y;
    const int dy_a = oy0 + strip * rows_per_strip;
    const int dy_b = min(dy_a + rows_per_strip, oy0 + ny);
    const int r_lo = __ldg(ay.start + dy_a);
    const int nrows = __ldg(ay.start + dy_b - 1) + __ldg(ay.cnt + dy_b - 1) - r_lo;
    const uint8_t* gbase = src + f * frame_stride + xb0;
    const int delta = static_cast<int>(reinterpret_cast<uintptr_t>(gbase) & 15);

    if (tid == 0) {
        for (int s = 0; s < AH_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], ncons >> 5); }
        fence_mbar_init();
    }
    // row program of the strip: one thread per output row fills the entries of the rows it taps (a row shared by two
    // output rows gets beta_cur/finish from the upper one and beta_next from the lower one: disjoint fields)
    for (int i = tid; i < nrows; i += blockDim.x) rinfo[i].beta_next = 0.f, rinfo[i].finish = 0, rinfo[i].pad = 0;
    __syncthreads();
    for (int d = tid; d < dy_b - dy_a; d += blockDim.x) {
        const int dy = dy_a + d;
        const int sy0 = __ldg(ay.start + dy), cy = __ldg(ay.cnt + dy);
        const bool shared_first = d > 0 && sy0 == __ldg(ay.start + dy - 1) + __ldg(ay.cnt + dy - 1) - 1;
        for (int j = 0; j < cy; ++j) {
            const float beta = __ldg(ay.wf + dy * ay.stride + j);
            AhRowInfo& ri = rinfo[sy0 + j - r_lo];
            // pad: integer weights (INTX) -- low half into the current output row, high half into the next one; the two
            // writers of a straddled row touch different halves, hence the atomic OR
            const int ib = INTX ? __float2int_rn(beta * static_cast<float>(ip.dy)) : 0;
            if (j == 0 && shared_first) { ri.beta_next = beta; if (INTX) atomicOr(&ri.pad, ib << 16); }
            else { ri.beta_cur = beta; if (INTX) atomicOr(&ri.pad, ib); }
            if (j == cy - 1) ri.finish = 1;
        }
    }
    __syncthreads();

    if (tid >= ncons) {
        // ------------------------------------------------------------------ producer warp
        if (tid == ncons) {
            const uint8_t* g = gbase - delta + static_cast<int64_t>(r_lo) * row_stride;
            for (int i = 0; i < nrows; ++i, g += row_stride) {
                const int s = i % AH_NSTAGE;
                if (i >= AH_NSTAGE) mbar_wait(&empty_bar[s], ((i / AH_NSTAGE) - 1) & 1, 11);
                mbar_arrive_expect_tx(&full_bar[s], seg_bytes);
                bulk_load_1d(ring + s * stage_bytes, g, seg_bytes, &full_bar[s]);
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    bool a_active[PX];
    int xcol[PX];
    uint32_t sh[PX], my_ring[PX];
    uint32_t qw[PX][3][4];          // INTX: per channel c and realigned word k, the four byte weights
    float wx[5];                    // fp32 path (PX == 1)
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        a_active[p] = tid + p * ncons < nx;
        xcol[p] = min(tid + p * ncons, nx - 1);
        const int dx = ox0 + xcol[p];
        const int cx = __ldg(ax.cnt + dx);
        const float* wxp = ax.wf + dx * ax.stride;
#pragma unroll
        for (int i = 0; i < 5; ++i) wx[i] = i < cx ? __ldg(wxp + i) : 0.f;
        if (INTX) {
            int ixw[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) ixw[i] = __float2int_rn(wx[i] * static_cast<float>(ip.dx));
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t q = 0;
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int n = 4 * k + m;   // byte n of the row window is tap n/3, channel n%3
                        if (n < 15 && n % 3 == c) q |= static_cast<uint32_t>(ixw[n / 3]) << (8 * m);
                    }
                    qw[p][c][k] = q;
                }
        }
        const int boff = __ldg(ax.start + dx) * 3 - xb0 + delta;      // first byte of this column inside a ring stage
        sh[p] = static_cast<uint32_t>(boff & 3) * 8;
        my_ring[p] = smem_u32(ring) + (boff & ~3);
    }
    // fp32 path constants (PX == 1: wx[] holds column 0)
    f32x2_t w01[5], c01[5], w2[2], c2[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        w01[i] = f2_pack(wx[i], wx[i]);
        c01[i] = f2_pack(-8388608.0f * wx[i], -8388608.0f * wx[i]);
    }
    w2[0] = f2_pack(wx[0], wx[1]); c2[0] = f2_pack(-8388608.0f * wx[0], -8388608.0f * wx[1]);
    w2[1] = f2_pack(wx[2], wx[3]); c2[1] = f2_pack(-8388608.0f * wx[2], -8388608.0f * wx[3]);
    const float w24 = wx[4], c24 = -8388608.0f * wx[4];
    const f32x2_t nz2 = f2_pack(negzero, negzero);
    auto magic = [](const uint32_t (&u)[4], int n) {   // 2^23 + byte n
        return __uint_as_float(__byte_perm(u[n >> 2], 0x4B000000u, 0x7650u | (n & 3)));
    };
    // horizontal Pillow pass: output columns left + tid (+ ncons), at most 7 taps each
    bool b_active[PX];
    int bk[PX][7], aword[PX], tcol[PX];
    uint32_t bsh[PX];
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        b_active[p] = tid + p * ncons < S;
        tcol[p] = min(tid + p * ncons, S - 1);
        const int ox = left + tcol[p];
        const int blo = __ldg(bx.start + ox), bcnt = __ldg(bx.cnt + ox);
#pragma unroll
        for (int i = 0; i < 7; ++i) bk[p][i] = i < bcnt ? __ldg(bx.wi + ox * bx.stride + i) : 0;
        const int aoff = (blo - ox0) * 3;                              // first byte inside a parked area row
        bsh[p] = static_cast<uint32_t>(aoff & 3) * 8;
        aword[p] = aoff & ~3;
    }

    const int lane = tid & 31;
    f32x2_t s01 = 0;     // fp32: 0 + beta*buf == beta*buf exactly (all terms are >= +0)
    float s2 = 0.f;
    int nacc[PX][3];     // INTX vertical accumulators
#pragma unroll
    for (int p = 0; p < PX; ++p) nacc[p][0] = nacc[p][1] = nacc[p][2] = 0;
    int par = 0;         // parity of the parked-row double buffer
    uint8_t* out_row = mid2 + f * mid2_frame_stride + static_cast<int64_t>(dy_a - oy0) * S * 3;
    // running shared-memory addresses (32-bit) of the stage being consumed: no per-row multiplies or cvta
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
    uint32_t soff = 0, fb = full0, eb = empty0, phase = 0, ria = smem_u32(rinfo);
    int s = 0;
    for (int i = 0; i < nrows; ++i) {
        {
            uint32_t ok;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(fb), "r"(phase) : "memory");
            if (!ok) {                                 // slow path: bounded spin (a protocol bug must trap, not hang)
                uint32_t spins = 0;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(fb), "r"(phase) : "memory");
                    if (!ok && ++spins > (1u << 24)) __trap();
                } while (!ok);
            }
        }
        uint32_t u[PX][4];
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            uint32_t w[5];
            asm volatile("ld.shared.u32 %0, [%5];\n\tld.shared.u32 %1, [%5+4];\n\tld.shared.u32 %2, [%5+8];\n\t"
                         "ld.shared.u32 %3, [%5+12];\n\tld.shared.u32 %4, [%5+16];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]) : "r"(my_ring[p] + soff));
#pragma unroll
            for (int k = 0; k < 4; ++k) u[p][k] = __funnelshift_r(w[k], w[k + 1], sh[p]);
        }
        float4 ri;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ri.x), "=f"(ri.y), "=f"(ri.z), "=f"(ri.w) : "r"(ria));
        ria += 16;
        __syncwarp();
        if (lane == 0)                                // this warp holds its bytes in registers now
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(eb) : "memory");
        soff += stage_bytes; fb += 8; eb += 8;
        if (++s == AH_NSTAGE) { s = 0; soff = 0; fb = full0; eb = empty0; phase ^= 1; }
        f32x2_t b01 = 0;
        float b2 = 0.f;
        int hsum[PX][3];
        if (INTX) {
            const int iyc = __float_as_int(ri.w) & 0xffff;
#pragma unroll
            for (int p = 0; p < PX; ++p) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    uint32_t hh = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) hh = __dp4a(u[p][k], qw[p][c][k], hh);
                    hsum[p][c] = static_cast<int>(hh);
                    nacc[p][c] += iyc * hsum[p][c];
                }
            }
        } else {
            b01 = f2_fma(f2_pack(magic(u[0], 0), magic(u[0], 1)), w01[0], c01[0]);
#pragma unroll
            for (int q = 1; q < 5; ++q)
                b01 = f2_add(b01, f2_fma(f2_pack(magic(u[0], q * 3), magic(u[0], q * 3 + 1)), w01[q], c01[q]));
            float p0, p1, p2, p3;
            f2_unpack(f2_fma(f2_pack(magic(u[0], 2), magic(u[0], 5)), w2[0], c2[0]), p0, p1);
            f2_unpack(f2_fma(f2_pack(magic(u[0], 8), magic(u[0], 11)), w2[1], c2[1]), p2, p3);
            const float p4 = __fmaf_rn(magic(u[0], 14), w24, c24);
            b2 = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(p0, p1), p2), p3), p4);
            s01 = f2_add(s01, f2_fma(f2_pack(ri.x, ri.x), b01, nz2));
            s2 = __fadd_rn(s2, __fmaf_rn(ri.x, b2, negzero));
        }
        if (__float_as_int(ri.z) != 0) {              // uniform over the CTA: an area-output row is complete
            uint8_t* ar = arow + par * arow_pitch;
            par ^= 1;
            if (INTX) {
                const int iyn = __float_as_int(ri.w) >> 16;
#pragma unroll
                for (int p = 0; p < PX; ++p) {
                    if (a_active[p]) {               // rint(N / D) == floor((2N + D) / 2D): no ties for odd D
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            ar[xcol[p] * 3 + c] = static_cast<uint8_t>(__umulhi(2u * nacc[p][c] + ip.d, ip.div_mul) >> ip.div_shift);
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) nacc[p][c] = iyn * hsum[p][c];
                }
            } else {
                if (a_active[0]) {
                    float s0, s1;
                    f2_unpack(s01, s0, s1);
                    ar[xcol[0] * 3 + 0] = static_cast<uint8_t>(min(max(__float2int_rn(s0), 0), 255));
                    ar[xcol[0] * 3 + 1] = static_cast<uint8_t>(min(max(__float2int_rn(s1), 0), 255));
                    ar[xcol[0] * 3 + 2] = static_cast<uint8_t>(min(max(__float2int_rn(s2), 0), 255));
                }
                // the straddling row opens the next output row (beta_next is 0 when there is none: 0*b = +0)
                s01 = f2_fma(f2_pack(ri.y, ri.y), b01, nz2);
                s2 = __fmaf_rn(ri.y, b2, negzero);
            }
            named_bar_sync(1, ncons);            // parked row complete (double buffered: one barrier per row)
#pragma unroll
            for (int p = 0; p < PX; ++p) {
                if (b_active[p]) {
                    const uint32_t* wp = reinterpret_cast<const uint32_t*>(ar + aword[p]);
                    uint32_t bw[7], bu[6];
#pragma unroll
                    for (int k = 0; k < 7; ++k) bw[k] = wp[k];
#pragma unroll
                    for (int k = 0; k < 6; ++k) bu[k] = __funnelshift_r(bw[k], bw[k + 1], bsh[p]);
                    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
#pragma unroll
                    for (int q = 0; q < 7; ++q) {
                        a0 += bk[p][q] * byte_as_int<6>(bu, q * 3);
                        a1 += bk[p][q] * byte_as_int<6>(bu, q * 3 + 1);
                        a2 += bk[p][q] * byte_as_int<6>(bu, q * 3 + 2);
                    }
                    uint8_t* o = out_row + tcol[p] * 3;
                    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
                }
            }
            out_row += S * 3;
        }
    }
    (void)max_rows;
}

// A+B fused, VERTICAL-FIRST integer form (INTX geometries; the 1080p / 720p bench path).  Same CTA decomposition,
// producer warp and bulk-copy ring as area_hpass_bulk_kernel, same integers out -- but the consumers accumulate down
// the rows BEFORE they combine across a row:
//   * per source row, consumer t owns the 16 bytes [16t, 16t+16) of the ring stage: one LDS.128, the even and the odd
//     bytes split into 16-bit lanes (two masks per word) and multiplied by the row's integer weight as packed pairs
//     (one IMAD per two bytes; a lane never exceeds 255 * Dy < 2^16).  That is 4 instructions per source word instead
//     of the ~8.5 (byte dot products + realignment) of the horizontal-first form, and the per-row bookkeeping is paid
//     once per 16 bytes;
//   * when an area-output row completes, the lanes are parked in shared memory (even-byte stream L, odd-byte stream
//     H), and thread x combines the 15 sums under area column x with IDP.2A (two 16-bit sums x two byte weights per
//     instruction), divides exactly as the horizontal-first form and parks the uint8 area row;
//   * Pillow's horizontal pass then runs on the PREVIOUS parked row (one CTA barrier per area row, everything double
//     buffered).
// N = sum_y sum_x iy*ix*byte is the same integer in either order, so the result is bit-identical by construction.
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// VW: 16-byte slices of a stage row per consumer; PA / PB: area / Pillow output columns per consumer (thread t owns
// slices t + v*ncons, area columns t + p*ncons, Pillow columns t + p*ncons).  1080p and 720p run <2, 3, 2> with 128
// consumers: fewer, fatter threads pay the per-row pipeline bookkeeping once per 32 bytes.
// (Re-reading the per-column constants -- packed IDP.2A weights, Pillow coefficients -- from their L1-resident tables
// once per area row instead of holding them in ~38 registers buys a fourth CTA per SM but measured 15 % slower.)
// NST: stages (source rows) of the bulk-copy ring, a power of two.
//
// NV12 = true (the decoder-output frame feed): the source is a 4:2:0 frame as a hardware decoder leaves
// it -- a Y plane and an interleaved half-resolution UV plane -- instead of packed RGB.  The producer streams, per
// source row, the window's luma bytes and the chroma bytes of row r/2 into the stage (1.5 B/px of HBM traffic instead
// of 3); consumer t converts ITS 8 pixels to the 24 RGB bytes OpenCV's COLOR_YUV2RGB_NV12 would produce (BT.601
// limited range, 20-bit fixed point, bit-exact: oracle/nv12_ref.py) in registers and from there on runs the very same
// integer pipeline on them.  RGB never exists in HBM.  One slice = 8 px = 24 bytes = 6 words (RGB source: 16 bytes).
__device__ __forceinline__ uint32_t pack_sat_u8(int hi, int lo, uint32_t upper) {
    uint32_t d;     // d = sat_u8(lo) | sat_u8(hi) << 8 | upper << 16
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(upper));
    return d;
}
// 8 luma bytes + 4 (U, V) pairs -> 24 RGB bytes in memory order (R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3 | ...)
__device__ __forceinline__ void nv12_convert8(uint2 yy, uint2 uv, uint32_t (&w)[6]) {
    constexpr int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527;
    constexpr int RND = 1 << 19;
    const uint32_t ys[2] = {__vsubus4(yy.x, 0x10101010u), __vsubus4(yy.y, 0x10101010u)};    // max(0, Y - 16), 4 at a time
    int c[8][3];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t word = p < 2 ? uv.x : uv.y;
        const int u = static_cast<int>(__byte_perm(word, 0u, 0x4440u | (2 * (p & 1))));
        const int v = static_cast<int>(__byte_perm(word, 0u, 0x4440u | (2 * (p & 1) + 1)));
        const int ruv = v * CVR + (RND - 128 * CVR);
        const int guv = u * CUG + (v * CVG + (RND - 128 * CVG - 128 * CUG));
        const int buv = u * CUB + (RND - 128 * CUB);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int px = 2 * p + e;
            const int y = static_cast<int>(__byte_perm(ys[px >> 2], 0u, 0x4440u | (px & 3)));
            // (the shift as IMAD.HI -- high word of x * 2^12, FMA pipe instead of the ALU pipe that the byte permutes keep
            // at 58 % -- measured SLOWER: 2.36 vs 2.0 ms per 1024 frames)
            c[px][0] = (y * CY + ruv) >> 20;
            c[px][1] = (y * CY + guv) >> 20;
            c[px][2] = (y * CY + buv) >> 20;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {       // word k holds RGB bytes 4k .. 4k+3; byte n is pixel n/3, channel n%3
        const int n = 4 * k;
        w[k] = pack_sat_u8(c[(n + 1) / 3][(n + 1) % 3], c[n / 3][n % 3],
                           pack_sat_u8(c[(n + 3) / 3][(n + 3) % 3], c[(n + 2) / 3][(n + 2) % 3], 0u));
    }
}
template <int MAXT, int MINB, int VW, int PA, int PB, int NST, bool NV12>
__global__ void __launch_bounds__(MAXT, MINB)
area_hpass_vfirst_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride,
                         const uint8_t* __restrict__ src_uv, int64_t uv_frame_stride, int segpx,
                         uint8_t* __restrict__ mid2, int64_t mid2_frame_stride, int oy0, int ny, int ox0, int nx,
                         int rows_per_strip, int xb0, int seg_bytes, int stage_bytes, int arow_pitch, int vpitch,
                         int left, int S, DevTaps ax, DevTaps ay, DevTaps bx, AhIntParams ip,
                         const uint32_t* __restrict__ aq /*[area columns][8] packed weights, see build_plan*/,
                         int nstrips, int nitems, int max_rows,
                         const AhRowInfo* __restrict__ strip_rinfo /*[nstrips][max_rows] row programs, host built*/,
                         const int2* __restrict__ strip_meta /*[nstrips] (first source row, source rows)*/,
                         int null_consumers /*probe: 1 = consumers only drain the ring (load path alone), 2 = no loads (consumers alone)*/) {
    // PERSISTENT: a CTA walks work items (strip, frame) = blockIdx.x, + gridDim.x, ...  The per-thread constants (packed
    // area weights, Pillow coefficients, stream offsets), the barriers and the zeroed slack are set up once per CTA,
    // the row program of every strip comes ready-made from a host-built table (double buffered in shared memory), and
    // the bulk-copy ring keeps running across items: the producer streams the first rows of the next item while the
    // consumers finish the current one.
    constexpr int WPS = NV12 ? 6 : 4;        // words per slice
    extern __shared__ __align__(128) uint8_t ah_smem[];
    const int ncons = blockDim.x - 32;                 // consumer threads; the last warp is the producer
    const int tid = threadIdx.x;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ah_smem);
    uint64_t* empty_bar = full_bar + NST;
    uint8_t* ring = ah_smem + 2 * NST * sizeof(uint64_t);
    uint8_t* vbuf = ring + NST * stage_bytes;     // [2 parities][L | H] streams, vpitch bytes each
    uint8_t* arow = vbuf + 4 * vpitch;                  // [2] parked uint8 area rows
    AhRowInfo* rinfo = reinterpret_cast<AhRowInfo*>(arow + 2 * arow_pitch);

    const uint8_t* gbase0 = src + xb0;
    // (RGB: row and frame pitches are multiples of 16 bytes, so the misalignment of a row segment is the same for every
    // row of every frame; NV12: the launcher guarantees 16-byte aligned planes, pitches and window start -> delta = 0)
    const int delta = NV12 ? 0 : static_cast<int>(reinterpret_cast<uintptr_t>(gbase0) & 15);

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], ncons >> 5); }
        fence_mbar_init();
    }
    // the slack words behind the streams are read (with weight 0, or shifted out) but never written per row
    for (int i = tid; i < vpitch; i += blockDim.x) reinterpret_cast<uint32_t*>(vbuf)[i] = 0u;
    __syncthreads();

    if (tid >= ncons) {
        // ------------------------------------------------------------------ producer warp
        if (tid == ncons) {
            uint32_t it = 0;                              // rows issued so far: the ring runs across items
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int strip = item % nstrips;
                const int64_t f = item / nstrips;
                const int2 meta = __ldg(strip_meta + strip);
                const int r_lo = meta.x, nrows = meta.y;
                const uint8_t* g = gbase0 + f * frame_stride - delta + static_cast<int64_t>(r_lo) * row_stride;
                const uint8_t* guv = NV12 ? src_uv + f * uv_frame_stride + xb0 : nullptr;
                for (int i = 0; i < nrows; ++i, ++it, g += row_stride) {
                    const uint32_t s = it & (NST - 1);
                    if (it >= NST) mbar_wait(&empty_bar[s], ((it / NST) - 1) & 1, 11);
                    if (null_consumers == 2) { mbar_arrive(&full_bar[s]); continue; }   // probe: no loads, consumers at full speed
                    if (NV12) {     // luma row r and chroma row r / 2 of the window: [Y: segpx bytes | UV: segpx bytes]
                        mbar_arrive_expect_tx(&full_bar[s], 2 * segpx);
                        bulk_load_1d(ring + s * stage_bytes, g, segpx, &full_bar[s]);
                        bulk_load_1d(ring + s * stage_bytes + segpx, guv + static_cast<int64_t>((r_lo + i) >> 1) * row_stride, segpx,
                                     &full_bar[s]);
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], seg_bytes);
                        bulk_load_1d(ring + s * stage_bytes, g, seg_bytes, &full_bar[s]);
                    }
                }
            }
        }
        return;
    }
    // ---------------------------------------------------------------------- consumers
    // slices [16 (tid + v*ncons), +16) of every stage row (NV12: pixels [8 (tid + v*ncons), +8), i.e. RGB bytes
    // [24 (tid + v*ncons), +24)); a slice past the segment re-reads slice 0 and is never parked
    bool v_active[VW];
    uint32_t v_off[VW];     // byte offset of the slice in the (RGB) row == offset of its lanes in the parked streams
#pragma unroll
    for (int v = 0; v < VW; ++v) {
        v_active[v] = (tid + v * ncons) * (4 * WPS) < seg_bytes;
        v_off[v] = v_active[v] ? static_cast<uint32_t>(4 * WPS) * static_cast<uint32_t>(tid + v * ncons) : 0u;
    }
    // area columns tid + p*ncons: stream addresses, realignment shifts and the packed IDP.2A weights
    bool a_active[PA];
    int xcol[PA];
    uint32_t xoff[PA], yoff[PA], q[PA][8];
#pragma unroll
    for (int p = 0; p < PA; ++p) {
        a_active[p] = tid + p * ncons < nx;
        xcol[p] = min(tid + p * ncons, nx - 1);
        const int dx = ox0 + xcol[p];
        // first (RGB) byte of this column inside a stage row (NV12: xb0 is the first PIXEL of the converted window)
        const int s0 = NV12 ? (__ldg(ax.start + dx) - xb0) * 3 : __ldg(ax.start + dx) * 3 - xb0 + delta;
        const int parity = s0 & 1, ex = s0 >> 1, ey = ex + parity;    // X_i = V[s0 + 2i], Y_i = V[s0 + 1 + 2i]
        // bit 31 = "stream starts in the high lane" (funnel shift by 16); the shift count is taken as (word >> 27)
        xoff[p] = static_cast<uint32_t>((parity ? vpitch : 0) + 4 * (ex >> 1)) | (static_cast<uint32_t>(ex & 1) << 31);
        yoff[p] = static_cast<uint32_t>((parity ? 0 : vpitch) + 4 * (ey >> 1)) | (static_cast<uint32_t>(ey & 1) << 31);
#pragma unroll
        for (int k = 0; k < 8; ++k) q[p][k] = __ldg(aq + dx * 8 + k);
    }
    // horizontal Pillow pass: output columns left + tid + p*ncons, at most 7 taps
    bool b_active[PB];
    int bk[PB][7], tcol[PB];
    uint32_t bsh[PB], aword[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) {
        b_active[p] = tid + p * ncons < S;
        tcol[p] = min(tid + p * ncons, S - 1);
        const int ox = left + tcol[p];
        const int blo = __ldg(bx.start + ox);
#pragma unroll
        for (int i = 0; i < 7; ++i) bk[p][i] = i < bx.stride ? __ldg(bx.wi + ox * bx.stride + i) : 0;   // zero padded
        const int aoff = (blo - ox0) * 3;                              // first byte inside a parked area row
        bsh[p] = static_cast<uint32_t>(aoff & 3) * 8;
        aword[p] = static_cast<uint32_t>(aoff & ~3);
    }
    auto hpass_row = [&](const uint8_t* ar, uint8_t* orow) {
#pragma unroll
        for (int p = 0; p < PB; ++p) {
            if (b_active[p]) {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(ar + aword[p]);
                uint32_t bw[7], bu[6];
#pragma unroll
                for (int k = 0; k < 7; ++k) bw[k] = wp[k];
#pragma unroll
                for (int k = 0; k < 6; ++k) bu[k] = __funnelshift_r(bw[k], bw[k + 1], bsh[p]);
                int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    a0 += bk[p][t] * byte_as_int<6>(bu, t * 3);
                    a1 += bk[p][t] * byte_as_int<6>(bu, t * 3 + 1);
                    a2 += bk[p][t] * byte_as_int<6>(bu, t * 3 + 2);
                }
                uint8_t* o = orow + tcol[p] * 3;
                o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
            }
        }
    };

    const int lane = tid & 31;
    uint32_t aL[VW][WPS], aH[VW][WPS];
    int par = 0;         // parity of the parked-row / stream double buffers (runs across items)
    // 32-bit shared addresses, made opaque so that they stay in registers (ptxas otherwise re-derives them from
    // SR_CgaCtaId in every row: an S2R round trip on the critical path of the row loop)
    uint32_t full0 = smem_u32(full_bar), ring0 = smem_u32(ring), vb0 = smem_u32(vbuf), ri0 = smem_u32(rinfo);
    uint32_t sbytes = static_cast<uint32_t>(stage_bytes);
    asm volatile("" : "+r"(full0), "+r"(ring0), "+r"(vb0), "+r"(ri0), "+r"(sbytes));
    const uint32_t ri_bytes = static_cast<uint32_t>(max_rows) * static_cast<uint32_t>(sizeof(AhRowInfo));
    uint32_t it = 0;     // rows consumed so far: ring stage = it % NST, phase = (it / NST) & 1
    uint32_t ibuf = 0;   // row-program buffer of the current item
    for (int item = blockIdx.x; item < nitems; item += gridDim.x, ibuf ^= 1u) {
        const int strip = item % nstrips;
        const int64_t f = item / nstrips;
        const int nrows = __ldg(strip_meta + strip).y;
        const int dy_a = oy0 + strip * rows_per_strip;
        // this item's row program -> buffer ibuf.  Every consumer has left the previous item (which used the other
        // buffer) once it passes the barrier, and the item before that is long finished.
        for (int i = tid; i < nrows; i += ncons) {
            const uint4 e = __ldg(reinterpret_cast<const uint4*>(strip_rinfo + static_cast<size_t>(strip) * max_rows + i));
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ri0 + ibuf * ri_bytes + 16u * i), "r"(e.x), "r"(e.y), "r"(e.z), "r"(e.w) : "memory");
        }
        named_bar_sync(1, ncons);
        uint32_t ria = ri0 + ibuf * ri_bytes;
#pragma unroll
        for (int v = 0; v < VW; ++v)
#pragma unroll
            for (int k = 0; k < WPS; ++k) aL[v][k] = aH[v][k] = 0u;
        int nfin = 0;        // area rows of this item finished so far
        uint8_t* out_row = mid2 + f * mid2_frame_stride + static_cast<int64_t>(dy_a - oy0) * S * 3;
        for (int i = 0; i < nrows; ++i, ++it) {
            const uint32_t s = it & (NST - 1), phase = (it / NST) & 1u;
            const uint32_t fb = full0 + 8u * s, eb = fb + 8u * NST, soff = ring0 + s * sbytes;
            {
                uint32_t ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                             "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(fb), "r"(phase) : "memory");
                if (!ok) {                                 // slow path: bounded spin (a protocol bug must trap, not hang)
                    uint32_t spins = 0;
                    do {
                        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(fb), "r"(phase) : "memory");
                        if (!ok && ++spins > (1u << 24)) __trap();
                    } while (!ok);
                }
            }
            uint32_t w[VW][WPS];
            uint2 nvy[VW], nvc[VW];
    #pragma unroll
            for (int v = 0; v < VW; ++v) {
                if constexpr (NV12) {       // 8 luma bytes and their 4 (U, V) pairs; v_off = 24 * slice -> byte 8 * slice
                    const uint32_t o = v_off[v] / 3u;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(nvy[v].x), "=r"(nvy[v].y) : "r"(soff + o));
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(nvc[v].x), "=r"(nvc[v].y) : "r"(soff + static_cast<uint32_t>(segpx) + o));
                } else {
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w[v][0]), "=r"(w[v][1]), "=r"(w[v][2]), "=r"(w[v][3]) : "r"(soff + v_off[v]));
                }
            }
            uint32_t fin, wts;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(fin), "=r"(wts) : "r"(ria + 8));
            ria += 16;
            __syncwarp();
            if (lane == 0)                                // this warp holds its bytes in registers now
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(eb) : "memory");
            if constexpr (NV12) {
    #pragma unroll
                for (int v = 0; v < VW; ++v) nv12_convert8(nvy[v], nvc[v], w[v]);
            }
            if (null_consumers == 1) { aL[0][0] ^= w[0][0] ^ w[VW - 1][3]; continue; }
            const uint32_t iyc = wts & 0xffffu;
            uint32_t lo[VW][WPS], hi[VW][WPS];
    #pragma unroll
            for (int v = 0; v < VW; ++v)
    #pragma unroll
                for (int k = 0; k < WPS; ++k) {
                    lo[v][k] = __byte_perm(w[v][k], 0u, 0x4240);     // bytes 0 and 2 in 16-bit lanes
                    hi[v][k] = __byte_perm(w[v][k], 0u, 0x4341);     // bytes 1 and 3
                    aL[v][k] += iyc * lo[v][k];
                    aH[v][k] += iyc * hi[v][k];
                }
            if (fin != 0) {                               // uniform over the CTA: an area-output row is complete
                const uint32_t vb = vb0 + static_cast<uint32_t>(par) * 2u * static_cast<uint32_t>(vpitch);
    #pragma unroll
                for (int v = 0; v < VW; ++v) {
                    if (v_active[v]) {
                        if constexpr (NV12) {       // 24-byte slices: 8-byte aligned
    #pragma unroll
                            for (int k = 0; k < WPS; k += 2) {
                                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(vb + v_off[v] + 4u * k), "r"(aL[v][k]), "r"(aL[v][k + 1]) : "memory");
                                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(vb + vpitch + v_off[v] + 4u * k), "r"(aH[v][k]), "r"(aH[v][k + 1]) : "memory");
                            }
                        } else {
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(vb + v_off[v]), "r"(aL[v][0]), "r"(aL[v][1]), "r"(aL[v][2]), "r"(aL[v][3]) : "memory");
                            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(vb + vpitch + v_off[v]), "r"(aH[v][0]), "r"(aH[v][1]), "r"(aH[v][2]), "r"(aH[v][3]) : "memory");
                        }
                    }
                }
                const uint32_t iyn = wts >> 16;           // a straddling row opens the next output row
    #pragma unroll
                for (int v = 0; v < VW; ++v)
    #pragma unroll
                    for (int k = 0; k < WPS; ++k) { aL[v][k] = iyn * lo[v][k]; aH[v][k] = iyn * hi[v][k]; }
                named_bar_sync(1, ncons);            // streams of this row complete; the previous parked row complete
                uint8_t* ar = arow + par * arow_pitch;
    #pragma unroll
                for (int p = 0; p < PA; ++p) {
                    if (a_active[p]) {
                        const uint32_t xa = vb + (xoff[p] & 0x7fffffffu), ya = vb + (yoff[p] & 0x7fffffffu);
                        const uint32_t xsh = xoff[p] >> 27, ysh = yoff[p] >> 27;      // 0 or 16
                        uint32_t xr[5], yr[5], x[4], y[4];
    #pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(xr[k]) : "r"(xa + 4u * k));
                            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(yr[k]) : "r"(ya + 4u * k));
                        }
    #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            x[k] = __funnelshift_r(xr[k], xr[k + 1], xsh);
                            y[k] = __funnelshift_r(yr[k], yr[k + 1], ysh);
                        }
                        const uint32_t (&qq)[8] = q[p];
                        uint32_t n0 = dp2a_lo(x[0], qq[0], 0u);
                        n0 = dp2a_hi(x[1], qq[0], n0);
                        n0 = dp2a_lo(x[3], qq[1], n0);
                        n0 = dp2a_hi(y[0], qq[1], n0);
                        n0 = dp2a_lo(y[2], qq[2], n0);
                        uint32_t n1 = dp2a_hi(x[1], qq[2], 0u);
                        n1 = dp2a_lo(x[2], qq[3], n1);
                        n1 = dp2a_hi(y[0], qq[3], n1);
                        n1 = dp2a_lo(y[1], qq[4], n1);
                        n1 = dp2a_hi(y[3], qq[4], n1);
                        uint32_t n2 = dp2a_lo(x[0], qq[5], 0u);
                        n2 = dp2a_hi(x[2], qq[5], n2);
                        n2 = dp2a_lo(x[3], qq[6], n2);
                        n2 = dp2a_hi(y[1], qq[6], n2);
                        n2 = dp2a_lo(y[2], qq[7], n2);
                        // rint(N / D) == floor((2N + D) / 2D): no ties for odd D
                        uint8_t* o = ar + xcol[p] * 3;
                        o[0] = static_cast<uint8_t>(__umulhi(2u * n0 + ip.d, ip.div_mul) >> ip.div_shift);
                        o[1] = static_cast<uint8_t>(__umulhi(2u * n1 + ip.d, ip.div_mul) >> ip.div_shift);
                        o[2] = static_cast<uint8_t>(__umulhi(2u * n2 + ip.d, ip.div_mul) >> ip.div_shift);
                    }
                }
                if (nfin > 0) {
                    hpass_row(arow + (par ^ 1) * arow_pitch, out_row);
                    out_row += S * 3;
                }
                ++nfin;
                par ^= 1;
            }
        }
        if (nfin > 0) {
            named_bar_sync(1, ncons);                // the last parked row of the item is complete
            hpass_row(arow + (par ^ 1) * arow_pitch, out_row);
        }
    }
    if (null_consumers == 1 && aL[0][0] == 0x12345678u) mid2[0] = 1;   // keeps the probe's loads alive
}

// B: Pillow horizontal pass for output columns [ocol0, ocol0+S) on rows [0, ny) of the (possibly compacted)
// source whose column 0 is absolute column src_x0.  FAST: at most 7 taps (21 bytes) -> aligned word loads.
template <bool FAST>
__global__ void __launch_bounds__(256)
hpass_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, int src_x0,
             uint8_t* __restrict__ dst, int64_t dst_frame_stride, int ny, int ocol0, int S, DevTaps bx) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(ny * S)) return;
    const int y = id / S, x = id - y * S;
    const int64_t f = blockIdx.y;
    const int ox = ocol0 + x;
    const int lo = bx.start[ox], cnt = bx.cnt[ox];
    const int* k = bx.wi + ox * bx.stride;
    const uint8_t* p = src + f * frame_stride + static_cast<int64_t>(y) * row_stride + (lo - src_x0) * 3;
    int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
    if (FAST) {
        uint32_t u[6];
        load_bytes_aligned<6>(p, cnt * 3, u);
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int kk = i < cnt ? k[i] : 0;   // bytes past cnt may be junk: weight 0
            a0 += kk * byte_as_int<6>(u, i * 3);
            a1 += kk * byte_as_int<6>(u, i * 3 + 1);
            a2 += kk * byte_as_int<6>(u, i * 3 + 2);
        }
    } else {
        for (int i = 0; i < cnt; ++i) {
            const int kk = k[i];
            a0 += kk * p[i * 3]; a1 += kk * p[i * 3 + 1]; a2 += kk * p[i * 3 + 2];
        }
    }
    uint8_t* o = dst + f * dst_frame_stride + (y * S + x) * 3;
    o[0] = clip8(a0); o[1] = clip8(a1); o[2] = clip8(a2);
}

// B, four output pixels per thread (S % 4 == 0, 4-byte aligned rows of dst): the 12 result bytes leave as three 32-bit
// stores instead of twelve byte stores, and the four independent tap chains give the scheduler something to overlap.
// Same arithmetic per pixel as hpass_kernel.
template <bool FAST>
__global__ void __launch_bounds__(256)
hpass4_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, int src_x0,
              uint8_t* __restrict__ dst, int64_t dst_frame_stride, int ny, int ocol0, int S, DevTaps bx) {
    const int sq = S >> 2;
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(ny * sq)) return;
    const int y = id / sq, x = (id - y * sq) << 2;
    const int64_t f = blockIdx.y;
    const uint8_t* rowp = src + f * frame_stride + static_cast<int64_t>(y) * row_stride;
    uint32_t w[3] = {0u, 0u, 0u};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ox = ocol0 + x + q;
        const int lo = __ldg(bx.start + ox), cnt = __ldg(bx.cnt + ox);
        const int* k = bx.wi + ox * bx.stride;
        const uint8_t* p = rowp + (lo - src_x0) * 3;
        int a0 = 1 << 21, a1 = 1 << 21, a2 = 1 << 21;
        if (FAST) {
            uint32_t u[6];
            load_bytes_aligned<6>(p, cnt * 3, u);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int kk = i < cnt ? __ldg(k + i) : 0;   // bytes past cnt may be junk: weight 0
                a0 += kk * byte_as_int<6>(u, i * 3);
                a1 += kk * byte_as_int<6>(u, i * 3 + 1);
                a2 += kk * byte_as_int<6>(u, i * 3 + 2);
            }
        } else {
            for (int i = 0; i < cnt; ++i) {
                const int kk = __ldg(k + i);
                a0 += kk * p[i * 3]; a1 += kk * p[i * 3 + 1]; a2 += kk * p[i * 3 + 2];
            }
        }
        const uint32_t b[3] = {clip8(a0), clip8(a1), clip8(a2)};
#pragma unroll
        for (int c = 0; c < 3; ++c) w[(q * 3 + c) >> 2] |= b[c] << (8 * ((q * 3 + c) & 3));
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(dst + f * dst_frame_stride + (static_cast<int64_t>(y) * S + x) * 3);
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
}

// C: vertical pass (or plain crop) + ToTensor/Normalize lookup + store.  One thread per 8 output pixels of one
// output row (all 3 channels): 24 contiguous source bytes per tap, fetched as aligned words.  Output either bf16
// patch-major rows (col = c*P*P + y*P + x, patch_k padded) or fp32 CHW.
__global__ void __launch_bounds__(128)
vpass_store_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, int src_y0, int src_x0,
                   int has_v, int top, DevTaps cy, int S, int P, int grid, int patch_k,
                   const float* __restrict__ lut /*[3][256]*/, bf16* __restrict__ patches, float* __restrict__ chw, int bgr) {
    __shared__ float s_lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const int xchunks = (S + 7) >> 3;
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(S * xchunks)) return;
    const int oy = id / xchunks, xc = id - oy * xchunks;
    const int64_t f = blockIdx.y;
    const int x0 = xc << 3;
    const int npx = min(8, S - x0);
    uint8_t u8[24];
    const uint8_t* fbase = src + f * frame_stride + (x0 - src_x0) * 3;
    if (has_v) {
        const int o = top + oy;
        const int lo = cy.start[o], cnt = cy.cnt[o];
        const int* k = cy.wi + o * cy.stride;
        int acc[24];
#pragma unroll
        for (int j = 0; j < 24; ++j) acc[j] = 1 << 21;
        for (int i = 0; i < cnt; ++i) {
            uint32_t u[6];
            load_bytes_aligned<6>(fbase + static_cast<int64_t>(lo + i - src_y0) * row_stride, npx * 3, u);
            const int kk = k[i];
#pragma unroll
            for (int j = 0; j < 24; ++j) acc[j] += kk * byte_as_int<6>(u, j);
        }
#pragma unroll
        for (int j = 0; j < 24; ++j) u8[j] = clip8(acc[j]);
    } else {
        uint32_t u[6];
        load_bytes_aligned<6>(fbase + static_cast<int64_t>(top + oy - src_y0) * row_stride, npx * 3, u);
#pragma unroll
        for (int j = 0; j < 24; ++j) u8[j] = static_cast<uint8_t>(byte_as_int<6>(u, j));
    }
    if (patches) {
        const int py = oy / P, yy = oy - py * P;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int co = bgr ? 2 - c : c;     // output plane / normalisation constants of source channel c
            // the 8 pixels may straddle a patch boundary when P is not a multiple of 8 -> element stores
            const int px0 = x0 / P;
            const bool one_patch = (npx == 8) && ((x0 + 7) / P == px0) && (((x0 - px0 * P) & 7) == 0) && ((P & 7) == 0);
            if (one_patch) {
                uint4 o;
                o.x = pack_bf16x2(s_lut[co * 256 + u8[0 * 3 + c]], s_lut[co * 256 + u8[1 * 3 + c]]);
                o.y = pack_bf16x2(s_lut[co * 256 + u8[2 * 3 + c]], s_lut[co * 256 + u8[3 * 3 + c]]);
                o.z = pack_bf16x2(s_lut[co * 256 + u8[4 * 3 + c]], s_lut[co * 256 + u8[5 * 3 + c]]);
                o.w = pack_bf16x2(s_lut[co * 256 + u8[6 * 3 + c]], s_lut[co * 256 + u8[7 * 3 + c]]);
                const int64_t row = (f * grid + py) * grid + px0;
                *reinterpret_cast<uint4*>(patches + row * patch_k + co * P * P + yy * P + (x0 - px0 * P)) = o;
            } else if ((P & 1) == 0 && (npx & 1) == 0 && (patch_k & 1) == 0) {
                // even patch size (P = 14): an even-aligned pixel pair never straddles a patch and its element index is
                // even, so the row leaves as 4-byte stores
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    if (j >= npx) break;
                    const int x = x0 + j;
                    const int px = x / P, xx = x - px * P;
                    if (px < grid && py < grid) {
                        const int64_t row = (f * grid + py) * grid + px;
                        *reinterpret_cast<uint32_t*>(patches + row * patch_k + co * P * P + yy * P + xx) =
                            pack_bf16x2(s_lut[co * 256 + u8[j * 3 + c]], s_lut[co * 256 + u8[(j + 1) * 3 + c]]);
                    }
                }
            } else {
                for (int j = 0; j < npx; ++j) {
                    const int x = x0 + j;
                    const int px = x / P, xx = x - px * P;
                    if (px < grid && py < grid) {
                        const int64_t row = (f * grid + py) * grid + px;
                        patches[row * patch_k + co * P * P + yy * P + xx] = __float2bfloat16(s_lut[co * 256 + u8[j * 3 + c]]);
                    }
                }
            }
        }
    }
    if (chw) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int co = bgr ? 2 - c : c;
            float* o = chw + ((f * 3 + co) * S + oy) * S + x0;
            for (int j = 0; j < npx; ++j) o[j] = s_lut[co * 256 + u8[j * 3 + c]];
        }
    }
}

// C, tile form (the bench path: patch-major bf16 output, P and S multiples of 8, 8-byte aligned 24-byte pixel groups,
// at most 7 vertical taps).  A CTA owns `rows_per_block` output rows of one frame, so the LUT is staged once per
// 32 rows instead of once per 4; thread (ry, xc) walks rows ry, ry + blockDim/xchunks, ...  All tap rows of an item
// are fetched up front as 64-bit loads (3 per tap, up to 21 in flight per thread) before the fixed-point
// accumulation -- the arithmetic (order, rounding, clip) is exactly vpass_store_kernel's.
template <bool HAS_V>
__global__ void __launch_bounds__(256)
vpass_store_tile_kernel(const uint8_t* __restrict__ src, int64_t frame_stride, int64_t row_stride, int src_y0,
                        int xoff_bytes, int top, DevTaps cy, int S, int P, int grid, int patch_k, int rows_per_block,
                        const float* __restrict__ lut /*[3][256]*/, bf16* __restrict__ patches, int bgr) {
    __shared__ float s_lut[768];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = __ldg(lut + i);
    __syncthreads();
    const int xchunks = S >> 3;
    const int rows_per_pass = blockDim.x / xchunks;
    const int ry = threadIdx.x / xchunks, xc = threadIdx.x - ry * xchunks;
    if (ry >= rows_per_pass) return;
    const int64_t f = blockIdx.y;
    const int x0 = xc << 3;
    const int px0 = x0 / P, xx = x0 - px0 * P;
    const uint8_t* fbase = src + f * frame_stride + xoff_bytes + x0 * 3;
    const int oy_end = min(S, static_cast<int>(blockIdx.x + 1) * rows_per_block);
    for (int oy = blockIdx.x * rows_per_block + ry; oy < oy_end; oy += rows_per_pass) {
        uint32_t u8w[6];      // the 24 result bytes, packed
        if (HAS_V) {
            const int o = top + oy;
            const int lo = __ldg(cy.start + o), cnt = __ldg(cy.cnt + o);
            const int* k = cy.wi + o * cy.stride;
            // the rows of one warp (it spans at most a few output rows) rarely need all 7 taps: taps >= the warp's
            // largest count are skipped by a warp-uniform branch (their weight would be 0)
            const int wcnt = __reduce_max_sync(__activemask(), cnt);
            int kk[7];
            uint32_t w[7][6];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                if (i < 4 || i < wcnt) {
                    const int ii = i < cnt ? i : cnt - 1;      // rows past cnt: re-read the last one with weight 0
                    kk[i] = i < cnt ? __ldg(k + i) : 0;
                    const uint2* rp = reinterpret_cast<const uint2*>(fbase + static_cast<int64_t>(lo + ii - src_y0) * row_stride);
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint2 v = __ldg(rp + q);
                        w[i][2 * q] = v.x; w[i][2 * q + 1] = v.y;
                    }
                }
            }
            int acc[24];
#pragma unroll
            for (int j = 0; j < 24; ++j) acc[j] = 1 << 21;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                if (i < 4 || i < wcnt) {
#pragma unroll
                    for (int j = 0; j < 24; ++j) acc[j] += kk[i] * byte_as_int<6>(w[i], j);
                }
            }
#pragma unroll
            for (int q = 0; q < 6; ++q)
                u8w[q] = static_cast<uint32_t>(clip8(acc[4 * q])) | (static_cast<uint32_t>(clip8(acc[4 * q + 1])) << 8) |
                         (static_cast<uint32_t>(clip8(acc[4 * q + 2])) << 16) | (static_cast<uint32_t>(clip8(acc[4 * q + 3])) << 24);
        } else {
            const uint2* rp = reinterpret_cast<const uint2*>(fbase + static_cast<int64_t>(top + oy - src_y0) * row_stride);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint2 v = __ldg(rp + q);
                u8w[2 * q] = v.x; u8w[2 * q + 1] = v.y;
            }
        }
        const int py = oy / P, yy = oy - py * P;
        const int64_t row = (f * grid + py) * grid + px0;
        bf16* orow = patches + row * patch_k + yy * P + xx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int co = bgr ? 2 - c : c;     // output plane / normalisation constants of source channel c
            const float* l = s_lut + co * 256;
            uint4 o;
            o.x = pack_bf16x2(l[byte_as_int<6>(u8w, 0 * 3 + c)], l[byte_as_int<6>(u8w, 1 * 3 + c)]);
            o.y = pack_bf16x2(l[byte_as_int<6>(u8w, 2 * 3 + c)], l[byte_as_int<6>(u8w, 3 * 3 + c)]);
            o.z = pack_bf16x2(l[byte_as_int<6>(u8w, 4 * 3 + c)], l[byte_as_int<6>(u8w, 5 * 3 + c)]);
            o.w = pack_bf16x2(l[byte_as_int<6>(u8w, 6 * 3 + c)], l[byte_as_int<6>(u8w, 7 * 3 + c)]);
            *reinterpret_cast<uint4*>(orow + co * P * P) = o;
        }
    }
}

// zero the K padding columns of the patch rows (only when 3*P*P is not a multiple of 64, e.g. P = 14)
__global__ void zero_pad_kernel(bf16* __restrict__ patches, int64_t rows, int kk, int patch_k) {
    const int pad = patch_k - kk;
    const int64_t total = rows * pad;
    for (int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; id < total;
         id += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = id / pad;
        patches[r * patch_k + kk + (id - r * pad)] = __float2bfloat16(0.f);
    }
}

// ---------------------------------------------------------------------------------------------- host
// Smallest d <= 255 such that every weight of outputs [o0, o1) is an integer multiple of 1/d (and each output's
// weights sum to exactly d); 0 if there is none.
static int common_denominator(const AxisTaps& t, int o0, int o1) {
    for (int d = 1; d <= 255; ++d) {
        bool ok = true;
        for (int o = o0; o < o1 && ok; ++o) {
            int sum = 0;
            for (int i = 0; i < t.cnt[o] && ok; ++i) {
                const double v = static_cast<double>(t.wf[static_cast<size_t>(o) * t.stride + i]) * d;
                const double r = nearbyint(v);
                if (fabs(v - r) > 1e-4 || r < 0 || r > 255) ok = false;
                sum += static_cast<int>(r);
            }
            if (sum != d) ok = false;
        }
        if (ok) return d;
    }
    return 0;
}

// Row programs of the vertical-first area kernel for strips of `rows` area rows: per strip the first source row and the
// number of source rows, and per source row (AhRowInfo) its weights into the output row being accumulated / the next
// one (a straddling row feeds both), the integer forms of the two (pad: low / high half) and the "output row
// complete" flag.  Built once per (geometry, rows) on the host; the persistent CTAs copy a strip's program into shared
// memory per work item instead of deriving it from the tap tables.
static int get_strip_table(b200clip_handle* h, const Plan& p, int rows, Plan::StripTable* out) {
    auto it = p.strips.find(rows);
    if (it != p.strips.end()) { *out = it->second; return 0; }
    const AxisTaps& ay = p.h_ay;
    const int oy0 = p.ry0, ny = p.ry1 - p.ry0;
    const int nstrips = (ny + rows - 1) / rows;
    const int max_rows = rows * ay.stride;
    std::vector<AhRowInfo> prog(static_cast<size_t>(nstrips) * max_rows, AhRowInfo{0.f, 0.f, 0, 0});
    std::vector<int2> meta(nstrips);
    for (int s = 0; s < nstrips; ++s) {
        const int dy_a = oy0 + s * rows, dy_b = std::min(dy_a + rows, oy0 + ny);
        const int r_lo = ay.start[dy_a];
        const int nrows = ay.start[dy_b - 1] + ay.cnt[dy_b - 1] - r_lo;
        if (nrows > max_rows) return b200_fail(h, B200CLIP_E_SHAPE, "preprocess: strip of %d rows needs %d source rows", rows, nrows);
        meta[s] = make_int2(r_lo, nrows);
        AhRowInfo* e = &prog[static_cast<size_t>(s) * max_rows];
        for (int dy = dy_a; dy < dy_b; ++dy) {
            const int sy0 = ay.start[dy], cy = ay.cnt[dy];
            const bool shared_first = dy > dy_a && sy0 == ay.start[dy - 1] + ay.cnt[dy - 1] - 1;
            for (int j = 0; j < cy; ++j) {
                const float beta = ay.wf[static_cast<size_t>(dy) * ay.stride + j];
                AhRowInfo& ri = e[sy0 + j - r_lo];
                const int ib = static_cast<int>(lrintf(beta * static_cast<float>(p.a_dy)));
                if (j == 0 && shared_first) { ri.beta_next = beta; ri.pad |= ib << 16; }
                else { ri.beta_cur = beta; ri.pad |= ib; }
                if (j == cy - 1) ri.finish = 1;
            }
        }
    }
    Plan::StripTable t;
    void* d = nullptr;
    B200_CUDA(h, cudaMalloc(&d, prog.size() * sizeof(AhRowInfo)));
    h->allocs.push_back(d);
    B200_CUDA(h, cudaMemcpy(d, prog.data(), prog.size() * sizeof(AhRowInfo), cudaMemcpyHostToDevice));
    t.rinfo = d;
    B200_CUDA(h, cudaMalloc(&d, meta.size() * sizeof(int2)));
    h->allocs.push_back(d);
    B200_CUDA(h, cudaMemcpy(d, meta.data(), meta.size() * sizeof(int2), cudaMemcpyHostToDevice));
    t.meta = d;
    t.nstrips = nstrips; t.max_rows = max_rows;
    p.strips[rows] = t;
    *out = t;
    return 0;
}

#include "preprocess_mma.cuh"

// Fragment tables of area_hpass_mma_kernel (see preprocess_mma.cuh) for source windows whose first byte sits `delta`
// bytes behind a 16-byte boundary.  Every weight is placed at the (register, byte) position at which the MMA fragment
// layouts of m16n8k32 expect it.  Stage 1 reads its K window with 32-bit loads in the natural order (K index kk of K step
// s = byte kb + 32 s + kk of the ring row, kb a multiple of 4); stage 3 with 64-bit loads, so K is enumerated in the
// order those deliver the bytes: kk of step s, lane (g, t4) = byte kb + 32 s + 8 t4 + 4 (kk >= 16) + (kk & 3).
static int get_mma_tables(b200clip_handle* h, const Plan& p, int delta, Plan::MmaTables* out) {
    auto it = p.mma.find(delta);
    if (it != p.mma.end()) { *out = it->second; return 0; }
    Plan::MmaTables t;
    const AxisTaps &ax = p.h_ax, &ay = p.h_ay, &bx = p.h_bx;
    const int S = h->cfg.image_size;
    const int nx = p.rx1 - p.rx0, ny = p.ry1 - p.ry0;
    const int gstart = p.sx0 * 3 - delta;                     // source-row byte held by byte 0 of a ring row
    // smallest pitch >= need that is `rem` modulo 128: the 8 rows a warp reads at once (stage 1: 16 bytes per row and
    // LDS.32; stage 3: 32 bytes per row and half-warp of an LDS.64) then fall on distinct shared-memory banks
    auto pitch_mod = [](int need, int rem) { int v = (need + 127) / 128 * 128 + rem; return v - 128 >= need ? v - 128 : v; };
    t.seg = (delta + (p.sx1 - p.sx0) * 3 + 15) & ~15;
    t.pitch1 = pitch_mod(t.seg + 64, 16);
    t.pitchA = pitch_mod(nx + 64, 32);
    t.ntx = (nx + 4) / 5;
    t.npt = (S + 15) / 16;
    t.ngroups = (ny + 7) / 8;
    bool ok = p.a_int && p.a_dx <= 255 && 2 * p.a_dy <= 255 && gstart >= 0;
    std::vector<uint32_t> a1(static_cast<size_t>(t.ntx) * 2 * 32 * 4, 0u), b2(static_cast<size_t>(t.ngroups) * 32 * 2, 0u),
        ap(static_cast<size_t>(t.npt) * 3 * 2 * 32 * 4, 0u);
    std::vector<int> kb1(t.ntx, 0), kbp(t.npt, 0);
    std::vector<int2> gmeta(t.ngroups);
    auto put = [](uint32_t& word, int j, uint32_t byte) { word |= (byte & 0xffu) << (8 * j); };
    // stage 1: 5 area pixels (15 interleaved columns) per tile against the <= 64 source-row bytes under them
    for (int ti = 0; ti < t.ntx && ok; ++ti) {
        const int px0 = p.rx0 + 5 * ti, px1 = std::min(px0 + 5, p.rx1);
        int lastb = 0;
        for (int x = px0; x < px1; ++x) lastb = std::max(lastb, 3 * (ax.start[x] + ax.cnt[x]) - gstart);
        const int kb = (3 * ax.start[px0] - gstart) & ~3;
        if (kb < 0 || lastb - kb > 64 || kb + 64 > t.pitch1) { ok = false; break; }
        kb1[ti] = kb;
        for (int s = 0; s < 2; ++s)
            for (int lane = 0; lane < 32; ++lane)
                for (int r = 0; r < 4; ++r)
                    for (int j = 0; j < 4; ++j) {
                        const int g = lane >> 2, t4 = lane & 3, m = g + 8 * (r & 1);
                        const int sb = kb + 32 * s + 16 * (r >> 1) + 4 * t4 + j + gstart;    // byte of the source row
                        const int sx = sb / 3, sc = sb % 3, x = px0 + m / 3;
                        if (m >= 15 || x >= px1 || sc != m % 3 || sx < ax.start[x] || sx >= ax.start[x] + ax.cnt[x]) continue;
                        const long iw = lrintf(ax.wf[static_cast<size_t>(x) * ax.stride + (sx - ax.start[x])] * static_cast<float>(p.a_dx));
                        if (iw < 0 || iw > 255) ok = false;
                        put(a1[((static_cast<size_t>(ti) * 2 + s) * 32 + lane) * 4 + r], j, static_cast<uint32_t>(iw));
                    }
    }
    // stage 2: 8 area rows per group against its <= 32 source rows, in the order the repacked accumulators hold them
    for (int gi = 0; gi < t.ngroups && ok; ++gi) {
        const int y0 = p.ry0 + 8 * gi, y1 = std::min(y0 + 8, p.ry1);
        const int r_lo = ay.start[y0];
        int r_hi = r_lo;
        for (int y = y0; y < y1; ++y) r_hi = std::max(r_hi, ay.start[y] + ay.cnt[y]);
        if (r_hi - r_lo > 32) { ok = false; break; }
        gmeta[gi] = make_int2(r_lo, r_hi - r_lo);
        t.nb = std::max(t.nb, (r_hi - r_lo + 7) / 8);
        for (int lane = 0; lane < 32; ++lane)
            for (int hh = 0; hh < 2; ++hh)
                for (int j = 0; j < 4; ++j) {
                    const int n = lane >> 2, t4 = lane & 3, y = y0 + n;
                    const int sr = r_lo + 16 * hh + (j < 2 ? 2 * t4 + j : 8 + 2 * t4 + (j - 2));
                    if (y >= y1 || sr < ay.start[y] || sr >= ay.start[y] + ay.cnt[y]) continue;
                    const long iw = 2 * lrintf(ay.wf[static_cast<size_t>(y) * ay.stride + (sr - ay.start[y])] * static_cast<float>(p.a_dy));
                    if (iw < 0 || iw > 255) ok = false;
                    put(b2[(static_cast<size_t>(gi) * 32 + lane) * 2 + hh], j, static_cast<uint32_t>(iw));
                }
    }
    // stage 3: 16 Pillow output columns per tile against the <= 64 area columns under them, three byte planes of the
    // 22-bit coefficients (two's complement: w = p0 + 256 p1 + 65536 p2, p2 signed)
    for (int pt = 0; pt < t.npt && ok; ++pt) {
        const int o0 = p.left + 16 * pt, o1 = std::min(o0 + 16, p.left + S);
        int lasta = 0;
        for (int o = o0; o < o1; ++o) lasta = std::max(lasta, bx.start[o] + bx.cnt[o] - p.rx0);
        const int kb = (bx.start[o0] - p.rx0) & ~7;
        if (kb < 0 || lasta - kb > 64 || kb + 64 > t.pitchA) { ok = false; break; }
        kbp[pt] = kb;
        for (int pl = 0; pl < 3; ++pl)
            for (int s = 0; s < 2; ++s)
                for (int lane = 0; lane < 32; ++lane)
                    for (int r = 0; r < 4; ++r)
                        for (int j = 0; j < 4; ++j) {
                            const int g = lane >> 2, t4 = lane & 3, o = o0 + g + 8 * (r & 1);
                            const int acol = kb + 32 * s + 8 * t4 + 4 * (r >> 1) + j + p.rx0;   // column of the area image
                            if (o >= o1 || acol < bx.start[o] || acol >= bx.start[o] + bx.cnt[o]) continue;
                            const int wv = bx.wi[static_cast<size_t>(o) * bx.stride + (acol - bx.start[o])];
                            if (wv < -(1 << 23) || wv >= (1 << 23)) ok = false;
                            put(ap[(((static_cast<size_t>(pt) * 3 + pl) * 2 + s) * 32 + lane) * 4 + r], j,
                                static_cast<uint32_t>(wv >> (8 * pl)));
                        }
    }
    if (ok) {
        Plan& pm = const_cast<Plan&>(p);
        int rc = 0;
        t.a1 = upload(h, pm, a1, rc); t.kb1 = upload(h, pm, kb1, rc); t.b2 = upload(h, pm, b2, rc);
        t.gmeta = upload(h, pm, gmeta, rc); t.ap = upload(h, pm, ap, rc); t.kbp = upload(h, pm, kbp, rc);
        if (rc) return b200_fail(h, rc, "preprocess: uploading the IMMA fragment tables failed");
    }
    t.ok = ok;
    p.mma[delta] = t;
    *out = t;
    return 0;
}

// CTAs of a K1 area launch.  Default: one CTA per work item.  B200CLIP_K1_PERSISTENT=1: every SM filled to the kernel's
// occupancy and each CTA walking many items -- measured on one box back to back (profiles/r02g_k1_persistent_ab.txt): 5.07 /
// 4.96 ms per 3600 frames persistent vs 4.99 / 4.76 one-shot, i.e. the per-CTA prologue is already hidden by the block
// scheduler with 3 resident CTAs per SM; kept as a parity-tested launch form.
template <class K>
static int persistent_grid(b200clip_handle* h, K kern, int threads, size_t smem, int nitems, int* cached) {
    if (!b200_knobs().k1_persistent) return nitems;
    if (*cached <= 0) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm <= 0) {
            cudaGetLastError();
            per_sm = 2;
        }
        *cached = per_sm;
    }
    const int g = *cached * h->num_sms;
    return nitems < g ? nitems : g;
}

static int build_plan(b200clip_handle* h, int H, int W, int mode, Plan& p) {
    const int S = h->cfg.image_size;
    p.H = H; p.W = W; p.mode = mode;
    int rc = 0;
    // stage A geometry (reference mode only): memory_manager.py:304-312
    int h1 = H, w1 = W;
    if (mode == B200CLIP_RESIZE_REFERENCE) {
        const double sw = 512.0 / W, sh = 512.0 / H;
        double scale = sw < sh ? sw : sh;
        if (scale > 1.0) scale = 1.0;
        if (scale < 1.0) {
            w1 = static_cast<int>(W * scale);
            h1 = static_cast<int>(H * scale);
            p.has_a = true;
        }
    }
    if (h1 <= 0 || w1 <= 0) return b200_fail(h, B200CLIP_E_SHAPE, "preprocess: degenerate frame %dx%d", W, H);
    p.h1 = h1; p.w1 = w1;
    // torchvision Resize(S): shortest side -> S, long side int(S * long / short)
    if (w1 <= h1) { p.nw = S; p.nh = static_cast<int>(static_cast<double>(S) * h1 / w1); }
    else { p.nh = S; p.nw = static_cast<int>(static_cast<double>(S) * w1 / h1); }
    p.left = py_round_half_even((p.nw - S) / 2.0);
    p.top = py_round_half_even((p.nh - S) / 2.0);
    p.has_b = p.nw != w1;
    p.has_c = p.nh != h1;
    const bool bicubic = mode != B200CLIP_RESIZE_BILINEAR_AA;
    AxisTaps bx, cy, ax, ay;
    if (p.has_c) {
        cy = pillow_taps(h1, p.nh, bicubic);
        taps_range(cy, p.top, p.top + S, p.ry0, p.ry1);
        for (int o = p.top; o < p.top + S; ++o) p.c_max_cnt = cy.cnt[o] > p.c_max_cnt ? cy.cnt[o] : p.c_max_cnt;
    } else {
        p.ry0 = p.top; p.ry1 = p.top + S;
    }
    if (p.has_b) {
        bx = pillow_taps(w1, p.nw, bicubic);
        taps_range(bx, p.left, p.left + S, p.rx0, p.rx1);
        for (int c : bx.cnt) p.b_max_cnt = c > p.b_max_cnt ? c : p.b_max_cnt;
    } else {
        p.rx0 = p.left; p.rx1 = p.left + S;
    }
    if (p.has_a) {
        const double scx = 1.0 / (w1 / static_cast<double>(W)), scy = 1.0 / (h1 / static_cast<double>(H));
        const int ix = static_cast<int>(scx), iy = static_cast<int>(scy);
        p.a_fast = fabs(scx - ix) < 2.220446049250313e-16 && fabs(scy - iy) < 2.220446049250313e-16;
        p.a_fx = ix; p.a_fy = iy;
        if (!p.a_fast) {
            ax = area_taps(W, w1);
            ay = area_taps(H, h1);
            for (int c : ax.cnt) p.a_max_cx = c > p.a_max_cx ? c : p.a_max_cx;
            p.a_seq = true;
            for (int dy = 0; dy + 1 < h1; ++dy) {
                const int end = ay.start[dy] + ay.cnt[dy];
                if (ay.cnt[dy] < 2 || (ay.start[dy + 1] != end && ay.start[dy + 1] != end - 1)) p.a_seq = false;
            }
            if (ay.cnt[h1 - 1] < 2) p.a_seq = false;
            // integer-exact evaluation: weights on a 1/(dx*dy) lattice with odd dx*dy (no rounding ties), and the fp32
            // evaluation error bound (2x margin) below the distance 1/(2*dx*dy) to the nearest tie
            int max_cy = 0;
            for (int dy = p.ry0; dy < p.ry1; ++dy) max_cy = ay.cnt[dy] > max_cy ? ay.cnt[dy] : max_cy;
            const int ddx = common_denominator(ax, p.rx0, p.rx1), ddy = common_denominator(ay, p.ry0, p.ry1);
            if (ddx > 0 && ddy > 0 && ((ddx * ddy) & 1)) {
                const double D = static_cast<double>(ddx) * ddy;
                const double err = 2.0 * (p.a_max_cx + max_cy + 6) * 255.0 * ldexp(1.0, -24);
                const uint32_t d2 = 2u * ddx * ddy, vmax = 2u * 255u * ddx * ddy + ddx * ddy;
                if (err < 1.0 / (2.0 * D)) {
                    for (int sft = 0; sft <= 16 && !p.a_int; ++sft) {     // magic divide by 2D, checked exhaustively
                        const uint64_t m = ((uint64_t(1) << (32 + sft)) + d2 - 1) / d2;
                        if (m >> 32) break;
                        bool ok = true;
                        for (uint32_t v = 0; v <= vmax && ok; ++v)
                            ok = static_cast<uint32_t>((static_cast<uint64_t>(v) * m) >> (32 + sft)) == v / d2;
                        if (ok) { p.a_int = true; p.a_dx = ddx; p.a_dy = ddy; p.a_div_shift = sft; p.a_div_mul = static_cast<uint32_t>(m); }
                    }
                }
            }
        }
    }
    // source window read by the first stage
    p.sx0 = p.rx0; p.sx1 = p.rx1; p.sy0 = p.ry0; p.sy1 = p.ry1;
    if (p.has_a) {
        if (p.a_fast) {
            p.sx0 = p.rx0 * p.a_fx; p.sx1 = p.rx1 * p.a_fx; p.sy0 = p.ry0 * p.a_fy; p.sy1 = p.ry1 * p.a_fy;
        } else {
            taps_range(ax, p.rx0, p.rx1, p.sx0, p.sx1);
            taps_range(ay, p.ry0, p.ry1, p.sy0, p.sy1);
        }
    }
    p.ax = upload_taps(h, p, ax, rc);
    p.ay = upload_taps(h, p, ay, rc);
    p.h_ay = ay;
    p.h_ax = ax;
    p.h_bx = bx;
    p.bx = upload_taps(h, p, bx, rc);
    p.cy = upload_taps(h, p, cy, rc);
    if (p.a_int && p.a_max_cx <= 5 && p.a_dx <= 255) {
        // area_hpass_vfirst_kernel: per area column, the byte pairs (weight of the low 16-bit lane, weight of the high
        // lane) in the order its combine consumes them, two pairs per word (.lo / .hi form of IDP.2A)
        std::vector<uint32_t> aq(static_cast<size_t>(w1) * 8, 0u);
        auto pr = [](uint32_t lo, uint32_t hi) { return lo | (hi << 8); };
        for (int dx = 0; dx < w1; ++dx) {
            uint32_t iw[5] = {0, 0, 0, 0, 0};
            for (int i = 0; i < ax.cnt[dx] && i < 5; ++i)
                iw[i] = static_cast<uint32_t>(lrintf(ax.wf[static_cast<size_t>(dx) * ax.stride + i] * static_cast<float>(p.a_dx)));
            uint32_t* q = &aq[static_cast<size_t>(dx) * 8];
            q[0] = pr(iw[0], 0) | (pr(0, iw[2]) << 16);     // channel 0: x0, x1
            q[1] = pr(iw[4], 0) | (pr(0, iw[1]) << 16);     //            x3, y0
            q[2] = pr(iw[3], 0) | (pr(iw[1], 0) << 16);     //            y2 | channel 1: x1
            q[3] = pr(0, iw[3]) | (pr(iw[0], 0) << 16);     //            x2, y0
            q[4] = pr(0, iw[2]) | (pr(iw[4], 0) << 16);     //            y1, y3
            q[5] = pr(0, iw[0]) | (pr(iw[2], 0) << 16);     // channel 2: x0, x2
            q[6] = pr(0, iw[4]) | (pr(iw[1], 0) << 16);     //            x3, y1
            q[7] = pr(0, iw[3]);                            //            y2
        }
        p.aq = upload(h, p, aq, rc);
    }
    if (rc) return b200_fail(h, rc, "preprocess: uploading resize tables failed");
    p.mid1_per_frame = p.has_a ? static_cast<size_t>(p.ry1 - p.ry0) * (p.rx1 - p.rx0) * 3 : 0;
    p.mid2_per_frame = p.has_b ? static_cast<size_t>(p.ry1 - p.ry0) * S * 3 : 0;
    // 16-byte align the per-frame strides
    p.mid1_per_frame = (p.mid1_per_frame + 15) & ~size_t(15);
    p.mid2_per_frame = (p.mid2_per_frame + 15) & ~size_t(15);
    return 0;
}

void preprocess_free_plans(b200clip_handle* h) {
    // the device tables themselves are owned by h->allocs
    delete static_cast<std::map<uint64_t, Plan>*>(h->pre_plans);
    h->pre_plans = nullptr;
}

static float* get_lut(b200clip_handle* h) {
    if (h->pre_lut) return h->pre_lut;
    // ToTensor (/255) then Normalize ((x - mean) / std) in fp32, exactly as torch evaluates it
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
    const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    std::vector<float> t(768);
    for (int c = 0; c < 3; ++c)
        for (int u = 0; u < 256; ++u) {
            volatile float x = static_cast<float>(u) / 255.0f;
            volatile float d = x - mean[c];
            t[c * 256 + u] = d / stdv[c];
        }
    float* d = nullptr;
    if (cudaMalloc(&d, 768 * sizeof(float)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, t.data(), 768 * sizeof(float), cudaMemcpyHostToDevice);
    h->allocs.push_back(d);
    h->pre_lut = d;
    return d;
}

// Probe builds only (-DB200CLIP_PROBES): B200CLIP_AREA_NULL=1 consumers only drain the ring (load path alone), =2 no
// loads (consumers alone).  Results are INVALID in either mode, so a release build cannot reach them.
static int area_null_probe() {
#ifdef B200CLIP_PROBES
    static const int v = getenv("B200CLIP_AREA_NULL") ? atoi(getenv("B200CLIP_AREA_NULL")) : 0;
    return v;
#else
    return 0;
#endif
}

static unsigned grid_for(b200clip_handle* h, int64_t total, int threads) {
    int64_t b = (total + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(h->num_sms) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<unsigned>(b);
}

static int get_plan(b200clip_handle* h, int H, int W, int mode, const Plan** out) {
    if (mode != B200CLIP_RESIZE_REFERENCE && mode != B200CLIP_RESIZE_BILINEAR_AA && mode != B200CLIP_RESIZE_BICUBIC)
        return b200_fail(h, B200CLIP_E_ARG, "preprocess: unknown resize mode %d", mode);
    if (H <= 0 || W <= 0 || H >= (1 << 28) || W >= (1 << 28))
        return b200_fail(h, B200CLIP_E_ARG, "preprocess: bad frame size %dx%d", W, H);
    const uint64_t key = (static_cast<uint64_t>(H) << 34) | (static_cast<uint64_t>(W) << 4) | static_cast<uint64_t>(mode);
    auto& pl = plans(h);
    auto it = pl.find(key);
    if (it == pl.end()) {
        Plan p;
        int rc = build_plan(h, H, W, mode, p);
        if (rc) return rc;
        it = pl.emplace(key, std::move(p)).first;
    }
    *out = &it->second;
    return 0;
}

// The rectangle of the source frame that K1 reads for this geometry (everything outside it is cropped away by the
// transform): the host-frame path uploads only this window.
int preprocess_source_window(b200clip_handle* h, int H, int W, int mode, int* x0, int* x1, int* y0, int* y1) {
    const Plan* pp = nullptr;
    int rc = get_plan(h, H, W, mode & ~B200CLIP_INPUT_BGR, &pp);
    if (rc) return rc;
    *x0 = pp->sx0; *x1 = pp->sx1; *y0 = pp->sy0; *y1 = pp->sy1;
    return 0;
}

// Stage C (+ the K padding of the patch rows) on `cur`: the image after the horizontal pass (or the source itself when
// no pass ran), whose element (0, 0) sits at stage coordinates (cur_y0, cur_x0).
static int run_stage_c(b200clip_handle* h, const Plan& p, const uint8_t* cur, int64_t cur_fs, int64_t cur_rs, int cur_x0,
                       int cur_y0, int n, bf16* patches, float* chw, cudaStream_t st, int bgr = 0) {
    const int S = h->cfg.image_size, P = h->cfg.patch;
    float* lut = get_lut(h);
    if (!lut) return b200_fail(h, B200CLIP_E_NOMEM, "preprocess: LUT allocation failed");
    {
        // stage C reads output-column x0 at cur column (x0 + xoff - cur_x0): after B the crop is already applied
        const int src_x0 = p.has_b ? 0 : cur_x0 - p.left;  // so that (x0 - src_x0) = x0 + left - cur_x0
        dim3 grid((static_cast<unsigned>(S) * ((S + 7) >> 3) + 127) / 128, n);
        ProfScope psc(h, PROF_PRE_C, static_cast<double>(n) * ((p.ry1 - p.ry0) * S * 3.0 + (patches ? static_cast<double>(h->grid) * h->grid * h->patch_k * 2.0 : 3.0 * S * S * 4.0)), st);
        // tile form: bf16 patch output, 8-pixel groups never straddle a patch, every 24-byte group 8-byte aligned
        const bool no_tile = b200_knobs().vpass_generic;   // parity tests cover both forms
        const int xoff = -src_x0 * 3;
        const bool tile = !no_tile && patches && !chw && (P & 7) == 0 && (S & 7) == 0 && (S >> 3) <= 256 &&
                          (!p.has_c || p.c_max_cnt <= 7) && (cur_rs & 7) == 0 && (cur_fs & 7) == 0 &&
                          ((reinterpret_cast<uintptr_t>(cur) + xoff) & 7) == 0;
        if (tile) {
            const int xchunks = S >> 3;
            const int threads = xchunks * (256 / xchunks);
            const int rows_per_block = 32;
            dim3 tgrid((S + rows_per_block - 1) / rows_per_block, n);
            if (p.has_c)
                vpass_store_tile_kernel<true><<<tgrid, threads, 0, st>>>(cur, cur_fs, cur_rs, cur_y0, xoff, p.top, p.cy, S, P,
                                                                         h->grid, h->patch_k, rows_per_block, lut, patches, bgr);
            else
                vpass_store_tile_kernel<false><<<tgrid, threads, 0, st>>>(cur, cur_fs, cur_rs, cur_y0, xoff, p.top, p.cy, S, P,
                                                                          h->grid, h->patch_k, rows_per_block, lut, patches, bgr);
        } else {
            vpass_store_kernel<<<grid, 128, 0, st>>>(cur, cur_fs, cur_rs, cur_y0, src_x0, p.has_c ? 1 : 0, p.top, p.cy, S, P,
                                                     h->grid, h->patch_k, lut, patches, chw, bgr);
        }
        h->launches++;
    }
    if (patches && h->patch_k != 3 * P * P) {
        const int64_t rows = static_cast<int64_t>(n) * h->grid * h->grid;
        zero_pad_kernel<<<grid_for(h, rows * (h->patch_k - 3 * P * P), 256), 256, 0, st>>>(patches, rows, 3 * P * P,
                                                                                          h->patch_k);
        h->launches++;
    }
    B200_CUDA(h, cudaGetLastError());
    return 0;
}

int launch_preprocess(b200clip_handle* h, const uint8_t* frames, int n, int H, int W, int64_t frame_stride,
                      int64_t row_stride, int mode, bf16* patches, float* chw, cudaStream_t st) {
    if (n <= 0) return 0;
    // B200CLIP_INPUT_BGR: every stage before the final store treats the three bytes of a pixel alike, so the channel
    // swap of cv2.cvtColor(BGR2RGB) (frame_extractor.py:191) is applied there (run_stage_c)
    const int bgr = (mode & B200CLIP_INPUT_BGR) ? 1 : 0;
    mode &= ~B200CLIP_INPUT_BGR;
    const Plan* pp = nullptr;
    int prc = get_plan(h, H, W, mode, &pp);
    if (prc) return prc;
    const Plan& p = *pp;
    // strides are checked against the window K1 reads, so that a compacted upload of that window is a valid input
    if (row_stride < static_cast<int64_t>(p.sx1 - p.sx0) * 3 || frame_stride < row_stride * (p.sy1 - p.sy0))
        return b200_fail(h, B200CLIP_E_ARG, "preprocess: bad frame geometry %dx%d strides %lld/%lld", W, H,
                         (long long)row_stride, (long long)frame_stride);
    const int S = h->cfg.image_size;
    // intermediates
    const size_t need = (p.mid1_per_frame + p.mid2_per_frame) * static_cast<size_t>(n);
    if (need > h->ws_pre_bytes) {
        B200_CUDA(h, cudaStreamSynchronize(st));
        if (h->ws_pre) cudaFree(h->ws_pre);
        h->ws_pre = nullptr; h->ws_pre_bytes = 0;
        B200_CUDA(h, cudaMalloc(&h->ws_pre, need));
        h->ws_pre_bytes = need;
    }
    uint8_t* mid1 = h->ws_pre;
    uint8_t* mid2 = h->ws_pre + p.mid1_per_frame * static_cast<size_t>(n);

    // algorithmic bytes: the whole frame in, the patch rows (or fp32 CHW) out
    ProfScope ps(h, PROF_PRE, static_cast<double>(n) * (static_cast<double>(H) * W * 3.0 +
                 (patches ? static_cast<double>(h->grid) * h->grid * h->patch_k * 2.0 : 0.0) +
                 (chw ? 3.0 * S * S * 4.0 : 0.0)), st);
    const uint8_t* cur = frames;
    int64_t cur_fs = frame_stride, cur_rs = row_stride;
    int cur_x0 = 0, cur_y0 = 0;  // absolute (stage-coordinate) position of cur's element (0,0)
    if (n > 65535) return b200_fail(h, B200CLIP_E_SHAPE, "preprocess: at most 65535 frames per call (got %d)", n);
    bool fused_ab = false;
    if (p.has_a && p.has_b && !p.a_fast && p.a_seq && p.a_max_cx <= 5 && p.b_max_cnt <= 7 && (row_stride & 15) == 0 &&
        (frame_stride & 15) == 0 && !b200_knobs().k1_unfused) {
        // A+B fused and bulk-copy fed: needs 16-byte aligned row segments that stay inside the row
        const int ny = p.ry1 - p.ry0, nx = p.rx1 - p.rx0;
        const int xb0 = p.sx0 * 3;
        const int delta = static_cast<int>((reinterpret_cast<uintptr_t>(frames) + xb0) & 15);
        const int seg = (delta + (p.sx1 - p.sx0) * 3 + 15) & ~15;
        const bool no_int = b200_knobs().area_fp32;    // parity tests cover every variant
        const bool no_px2 = b200_knobs().area_px1;
        const bool intx = p.a_int && !no_int;
        // two columns per thread when that still fits a 160-thread consumer group
        const int px = (intx && !no_px2 && (max(nx, S) + 1) / 2 <= 160) ? 2 : 1;
        const int ncons = ((max(nx, S) + px - 1) / px + 31) & ~31;
        // vertical-first integer form: one consumer per 16 bytes of the row window, two area columns per consumer
        const bool no_vfirst = b200_knobs().area_hfirst;
        // <2, 3, 2> on 128 consumers: 32 bytes of the row window, three area columns and two Pillow columns per thread
        // <2, 3, 2> on 128 consumers.  Measured alternatives, all within noise of 1.1-1.2 ms per 1024 frames: <1, 2, 1> on
        // 224 consumers at 3 or 4 CTAs per SM, a 16-stage ring, 24-288-row strips
        const int ncv = 128;
        const bool vfirst = intx && !no_vfirst && p.aq != nullptr && (seg >> 4) <= 2 * ncv && nx <= 3 * ncv && S <= 2 * ncv &&
                            p.a_dy * 255 < 65536 && p.a_dx <= 255 && xb0 - delta >= 0 && xb0 - delta + seg <= W * 3;
        // IMMA form (preprocess_mma.cuh): the three separable passes as integer tensor-core products, one persistent CTA
        // per SM.  Same integer arithmetic as the vertical-first kernel, so the same exactness conditions apply.
        // (the switches that pick among the CUDA-core kernels -- horizontal-first, one column per thread, persistent launch
        // form -- imply them)
        const bool want_mma = b200_knobs().area_mma && !b200_knobs().area_hfirst && !b200_knobs().area_px1 && !b200_knobs().k1_persistent;
        if (intx && want_mma && xb0 - delta >= 0) {
            Plan::MmaTables mt;
            if (int mrc = get_mma_tables(h, p, delta, &mt)) return mrc;
            constexpr int NCW = 15, TPW = 4;
            constexpr size_t kMmaSmem = 227 * 1024;      // the whole shared memory of an SM: one CTA per SM
            const size_t fixed = 256 + NBUF * 3 * (8 * static_cast<size_t>(mt.pitchA) + 8) + static_cast<size_t>(mt.npt) * 3 * 2 * 32 * 16 + NCW * 384;
            const size_t blockbytes = 8 * static_cast<size_t>(mt.pitch1);
            int nst = mt.ok && fixed < kMmaSmem ? static_cast<int>((kMmaSmem - fixed) / blockbytes) : 0;
            nst = nst > MMA_MAX_NST ? MMA_MAX_NST : nst;
            if (b200_knobs().k1_verbose)
                fprintf(stderr, "b200clip K1 %dx%d: IMMA tables ok=%d ntx=%d npt=%d groups=%d nb=%d seg=%d pitch1=%d pitchA=%d nst=%d\n", W, H,
                        (int)mt.ok, mt.ntx, mt.npt, mt.ngroups, mt.nb, mt.seg, mt.pitch1, mt.pitchA, nst);
            if (mt.ok && S % 16 == 0 && mt.ntx <= NCW * TPW && mt.nb >= 1 && mt.nb <= 4 && nst >= 3 && xb0 - delta + mt.seg <= W * 3) {
                int gps = 3;         // groups of 8 area rows per work item
                while (gps > 1 && static_cast<int64_t>(n) * ((mt.ngroups + gps - 1) / gps) < static_cast<int64_t>(h->num_sms) * 4) --gps;
                MmaParams mp{};
                mp.src = cur + xb0 - delta; mp.frame_stride = cur_fs; mp.row_stride = cur_rs;
                mp.mid2 = mid2; mp.mid2_frame_stride = static_cast<int64_t>(p.mid2_per_frame);
                mp.S = S; mp.ny = ny; mp.nx = nx; mp.seg = mt.seg; mp.pitch1 = mt.pitch1; mp.pitchA = mt.pitchA;
                mp.ntx = mt.ntx; mp.npt = mt.npt; mp.gps = gps; mp.nstrips = (mt.ngroups + gps - 1) / gps; mp.ngroups = mt.ngroups;
                mp.nitems = mp.nstrips * n; mp.nst = nst;
                mp.d = static_cast<uint32_t>(p.a_dx * p.a_dy); mp.div_mul = p.a_div_mul; mp.div_shift = p.a_div_shift;
                mp.a1 = static_cast<const uint4*>(mt.a1); mp.kb1 = static_cast<const int*>(mt.kb1);
                mp.b2 = static_cast<const uint2*>(mt.b2); mp.gmeta = static_cast<const int2*>(mt.gmeta);
                mp.ap = static_cast<const uint4*>(mt.ap); mp.kbp = static_cast<const int*>(mt.kbp);
                const size_t smem = fixed + nst * blockbytes;
                void (*kern)(const MmaParams) = mt.nb == 4 ? area_hpass_mma_kernel<NCW, TPW, 4>
                                              : mt.nb == 3 ? area_hpass_mma_kernel<NCW, TPW, 3>
                                              : mt.nb == 2 ? area_hpass_mma_kernel<NCW, TPW, 2> : area_hpass_mma_kernel<NCW, TPW, 1>;
                if (!(h->attr_done & ATTR_K1_MMA)) {
                    B200_CUDA(h, cudaFuncSetAttribute(area_hpass_mma_kernel<NCW, TPW, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMmaSmem)));
                    B200_CUDA(h, cudaFuncSetAttribute(area_hpass_mma_kernel<NCW, TPW, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMmaSmem)));
                    B200_CUDA(h, cudaFuncSetAttribute(area_hpass_mma_kernel<NCW, TPW, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMmaSmem)));
                    B200_CUDA(h, cudaFuncSetAttribute(area_hpass_mma_kernel<NCW, TPW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kMmaSmem)));
                    h->attr_done |= ATTR_K1_MMA;
                }
                ProfScope psa(h, PROF_PRE_A, static_cast<double>(n) * (static_cast<double>(p.sy1 - p.sy0) * (p.sx1 - p.sx0) * 3.0 + ny * S * 3.0), st);
                int fgrid = mp.nitems < h->num_sms ? mp.nitems : h->num_sms;
#ifdef B200CLIP_PROBES
                // probe (profiles/r02w_*): the kernel's time is nearly proportional to the CTAs it may use
                if (const char* e = getenv("B200CLIP_K1_CTAS")) { const int v = atoi(e); if (v > 0 && v < fgrid) fgrid = v; }
#endif
                kern<<<fgrid, (NCW + 1) * 32, smem, st>>>(mp);
                h->launches++;
                fused_ab = true;
                cur = mid2; cur_fs = p.mid2_per_frame; cur_rs = static_cast<int64_t>(S) * 3;
                cur_x0 = 0; cur_y0 = p.ry0;
            }
        }
        if (!fused_ab && vfirst) {
            // strip length: 24, 48, 96 and 288 rows measured within noise of each other on B200 (the per-CTA prologue
            // is ~10 % of the stall samples at 24 rows, but longer strips lose as much to the tail) -> keep 24
            int rows = 24;
            while (rows > 4 && static_cast<int64_t>(n) * ((ny + rows - 1) / rows) < static_cast<int64_t>(h->num_sms) * 6)
                rows = (rows + 1) / 2;
            const int stage_bytes = seg + 16, arow_pitch = (nx * 3 + 32 + 15) & ~15, vpitch = seg + 32;
            Plan::StripTable stt;
            if (int src = get_strip_table(h, p, rows, &stt)) return src;
            const int max_rows = stt.max_rows;
            const size_t smem = 2 * AH_NSTAGE * sizeof(uint64_t) + static_cast<size_t>(AH_NSTAGE) * stage_bytes +
                                4 * static_cast<size_t>(vpitch) + 2 * static_cast<size_t>(arow_pitch) +
                                2 * static_cast<size_t>(max_rows) * sizeof(AhRowInfo);
            if (smem <= 200 * 1024) {
                AhIntParams ip{p.a_dx, p.a_dy, p.a_dx * p.a_dy, p.a_div_shift, p.a_div_mul};
                auto kern = area_hpass_vfirst_kernel<160, 3, 2, 3, 2, AH_NSTAGE, false>;
                if (!(h->attr_done & ATTR_K1_VFIRST)) {
                    B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    // ~45 KB per CTA: without the maximum carve-out the driver's default split allows only 3 CTAs per SM
                    B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                    h->attr_done |= ATTR_K1_VFIRST;
                }
                ProfScope psa(h, PROF_PRE_A, static_cast<double>(n) * (static_cast<double>(p.sy1 - p.sy0) * (p.sx1 - p.sx0) * 3.0 + ny * S * 3.0), st);
                const int nitems = stt.nstrips * n;
                const int fgrid = persistent_grid(h, kern, ncv + 32, smem, nitems, &h->k1_ctas_per_sm[0]);
                kern<<<fgrid, ncv + 32, smem, st>>>(cur, cur_fs, cur_rs, nullptr, 0, 0, mid2, p.mid2_per_frame, p.ry0, ny, p.rx0, nx, rows, xb0,
                                                   seg, stage_bytes, arow_pitch, vpitch, p.left, S, p.ax, p.ay, p.bx, ip, p.aq,
                                                   stt.nstrips, nitems, max_rows, static_cast<const AhRowInfo*>(stt.rinfo),
                                                   static_cast<const int2*>(stt.meta), area_null_probe());
                h->launches++;
                fused_ab = true;
                cur = mid2; cur_fs = p.mid2_per_frame; cur_rs = static_cast<int64_t>(S) * 3;
                cur_x0 = 0; cur_y0 = p.ry0;
            }
        }
        if (!fused_ab && xb0 - delta >= 0 && xb0 - delta + seg <= W * 3 && ncons + 32 <= 576) {
            int rows = 24;
            while (rows > 4 && static_cast<int64_t>(n) * ((ny + rows - 1) / rows) < static_cast<int64_t>(h->num_sms) * 6)
                rows = (rows + 1) / 2;
            const int stage_bytes = seg + 16, arow_pitch = (nx * 3 + 32 + 15) & ~15;
            const int max_rows = rows * p.ay.stride;     // upper bound on the source rows of one strip
            const size_t smem = 2 * AH_NSTAGE * sizeof(uint64_t) + static_cast<size_t>(AH_NSTAGE) * stage_bytes +
                                2 * static_cast<size_t>(arow_pitch) + static_cast<size_t>(max_rows) * sizeof(AhRowInfo);
            if (smem <= 200 * 1024) {
                AhIntParams ip{p.a_dx, p.a_dy, p.a_dx * p.a_dy, p.a_div_shift, p.a_div_mul};
                auto kern = px == 2 ? area_hpass_bulk_kernel<192, 4, true, 2>
                          : ncons + 32 <= 352 ? (intx ? area_hpass_bulk_kernel<352, 3, true, 1> : area_hpass_bulk_kernel<352, 3, false, 1>)
                                              : (intx ? area_hpass_bulk_kernel<576, 2, true, 1> : area_hpass_bulk_kernel<576, 2, false, 1>);
                if (smem > 48 * 1024)
                    B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                ProfScope psa(h, PROF_PRE_A, static_cast<double>(n) * (static_cast<double>(p.sy1 - p.sy0) * (p.sx1 - p.sx0) * 3.0 + ny * S * 3.0), st);
                dim3 fgrid((ny + rows - 1) / rows, n);
                kern<<<fgrid, ncons + 32, smem, st>>>(cur, cur_fs, cur_rs, mid2, p.mid2_per_frame, p.ry0, ny, p.rx0, nx, rows,
                                                     max_rows, xb0, seg, stage_bytes, arow_pitch, p.left, S, p.ax, p.ay, p.bx, -0.0f, ip);
                h->launches++;
                fused_ab = true;
                cur = mid2; cur_fs = p.mid2_per_frame; cur_rs = static_cast<int64_t>(S) * 3;
                cur_x0 = 0; cur_y0 = p.ry0;
            }
        }
    }
    if (p.has_a && !fused_ab) {
        const int ny = p.ry1 - p.ry0, nx = p.rx1 - p.rx0;
        dim3 grid((static_cast<unsigned>(ny) * nx + 255) / 256, n);
        ProfScope psa(h, PROF_PRE_A, static_cast<double>(n) * (static_cast<double>(H) * (p.rx1 - p.rx0) * W / p.w1 * 3.0 + ny * nx * 3.0), st);
        const bool no_strip = b200_knobs().area_nostrip;   // parity tests cover both paths
        if (p.a_fast)
            area_fast_kernel<<<grid, 256, 0, st>>>(cur, cur_fs, cur_rs, mid1, p.mid1_per_frame, p.ry0, ny, p.rx0, nx,
                                                   p.a_fx, p.a_fy);
        else if (p.a_max_cx <= 5 && p.a_seq && (cur_rs & 3) == 0 && !no_strip) {
            // strips of output rows: long enough that the re-read boundary row is noise, short enough to fill the GPU
            int rows = 12;
            while (rows > 3 && static_cast<int64_t>(n) * nx * ((ny + rows - 1) / rows) < static_cast<int64_t>(h->num_sms) * 4096)
                rows = (rows + 1) / 2;
            const int nstrips = (ny + rows - 1) / rows;
            dim3 sgrid((static_cast<unsigned>(nstrips) * nx + 127) / 128, n);
            auto kern = area_strip_kernel<2>;     // source rows in flight: 1, 3 and 4 measured no better
            kern<<<sgrid, 128, 0, st>>>(cur, cur_fs, static_cast<int>(cur_rs >> 2), mid1, p.mid1_per_frame,
                                        p.ry0, ny, p.rx0, nx, rows, nstrips, p.sy1 - 1, p.ax, p.ay, -0.0f);
        } else if (p.a_max_cx <= 5)
            area_kernel_w5<<<grid, 256, 0, st>>>(cur, cur_fs, cur_rs, mid1, p.mid1_per_frame, p.ry0, ny, p.rx0, nx, p.ax,
                                                 p.ay);
        else
            area_kernel<<<grid, 256, 0, st>>>(cur, cur_fs, cur_rs, mid1, p.mid1_per_frame, p.ry0, ny, p.rx0, nx, p.ax,
                                              p.ay);
        h->launches++;
        cur = mid1; cur_fs = p.mid1_per_frame; cur_rs = static_cast<int64_t>(nx) * 3;
        cur_x0 = p.rx0; cur_y0 = p.ry0;
    }
    if (p.has_b && !fused_ab) {
        const int ny = p.ry1 - p.ry0;
        // rows [ry0, ry1) of cur: shift the base pointer so that row 0 of the kernel == row ry0
        const uint8_t* base = cur + static_cast<int64_t>(p.ry0 - cur_y0) * cur_rs;
        dim3 grid((static_cast<unsigned>(ny) * S + 255) / 256, n);
        ProfScope psb(h, PROF_PRE_B, static_cast<double>(n) * ny * (p.rx1 - p.rx0 + S) * 3.0, st);
        // four pixels per thread whenever the 12-byte groups of mid2 are word aligned (B200CLIP_HPASS_PX1=1: one pixel
        // per thread, the form the other shapes use)
        const bool px4 = !b200_knobs().hpass_px1 && (S & 3) == 0 && (p.mid2_per_frame & 3) == 0 && (reinterpret_cast<uintptr_t>(mid2) & 3) == 0;
        dim3 grid4((static_cast<unsigned>(ny) * (S >> 2) + 255) / 256, n);
        if (px4 && p.b_max_cnt <= 7)
            hpass4_kernel<true><<<grid4, 256, 0, st>>>(base, cur_fs, cur_rs, cur_x0, mid2, p.mid2_per_frame, ny, p.left, S, p.bx);
        else if (px4)
            hpass4_kernel<false><<<grid4, 256, 0, st>>>(base, cur_fs, cur_rs, cur_x0, mid2, p.mid2_per_frame, ny, p.left, S, p.bx);
        else if (p.b_max_cnt <= 7)
            hpass_kernel<true><<<grid, 256, 0, st>>>(base, cur_fs, cur_rs, cur_x0, mid2, p.mid2_per_frame, ny, p.left, S,
                                                     p.bx);
        else
            hpass_kernel<false><<<grid, 256, 0, st>>>(base, cur_fs, cur_rs, cur_x0, mid2, p.mid2_per_frame, ny, p.left,
                                                      S, p.bx);
        h->launches++;
        cur = mid2; cur_fs = p.mid2_per_frame; cur_rs = static_cast<int64_t>(S) * 3;
        cur_x0 = 0 /* column 0 of mid2 is output column `left`, handled below */; cur_y0 = p.ry0;
    }
    return run_stage_c(h, p, cur, cur_fs, cur_rs, cur_x0, cur_y0, n, patches, chw, st, bgr);
}


// ---------------------------------------------------------------------------------------------- NV12 frame feed
// Generic NV12 -> RGB conversion of the window K1 reads (every geometry / resize mode the fused kernel does not cover,
// and the B200CLIP_NV12_UNFUSED parity variant): one thread per horizontal pixel pair (one chroma sample), 6 bytes
// out.  Same arithmetic as nv12_convert8 (cv2 COLOR_YUV2RGB_NV12, oracle/nv12_ref.py).
__global__ void __launch_bounds__(256)
nv12_window_to_rgb_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ uvp, int64_t y_fs, int64_t uv_fs,
                          int64_t rs, uint8_t* __restrict__ dst, int64_t dst_fs, int64_t dst_pitch, int x0, int y0, int pairs,
                          int rows) {
    const unsigned id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= static_cast<unsigned>(pairs * rows)) return;
    const int r = id / pairs, c = id - r * pairs;
    const int64_t f = blockIdx.y;
    const int x = x0 + 2 * c, y = y0 + r;
    constexpr int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527;
    const uint8_t* yrow = yp + f * y_fs + static_cast<int64_t>(y) * rs + x;
    const uint8_t* uvrow = uvp + f * uv_fs + static_cast<int64_t>(y >> 1) * rs + x;
    const int u = static_cast<int>(uvrow[0]) - 128, v = static_cast<int>(uvrow[1]) - 128;
    const int ruv = (1 << 19) + CVR * v, guv = (1 << 19) + CVG * v + CUG * u, buv = (1 << 19) + CUB * u;
    uint8_t* o = dst + f * dst_fs + static_cast<int64_t>(r) * dst_pitch + c * 6;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int yy = max(0, static_cast<int>(yrow[e]) - 16) * CY;
        o[e * 3 + 0] = static_cast<uint8_t>(min(max((yy + ruv) >> 20, 0), 255));
        o[e * 3 + 1] = static_cast<uint8_t>(min(max((yy + guv) >> 20, 0), 255));
        o[e * 3 + 2] = static_cast<uint8_t>(min(max((yy + buv) >> 20, 0), 255));
    }
}

// Window of an NV12 frame K1 needs for this geometry: pixel columns [x0, x1) (x0 a multiple of 16, x1 of 16 or W) and
// rows [y0, y1); chroma rows [y0 / 2, (y1 + 1) / 2).  The host-frame path uploads exactly this.
int preprocess_source_window_nv12(b200clip_handle* h, int H, int W, int mode, int* x0, int* x1, int* y0, int* y1) {
    int rc = preprocess_source_window(h, H, W, mode, x0, x1, y0, y1);
    if (rc) return rc;
    *x0 &= ~15;
    *x1 = (*x1 + 15) & ~15;
    if (*x1 > W) *x1 = W;
    return 0;
}

// NV12 frames (Y plane at y, interleaved UV plane at uv, common row pitch rs) -> patch rows / CHW, bit-identical to
// cv2.cvtColor(COLOR_YUV2RGB_NV12) followed by launch_preprocess.
int launch_preprocess_nv12(b200clip_handle* h, const uint8_t* y, const uint8_t* uv, int n, int H, int W, int64_t y_fs,
                           int64_t uv_fs, int64_t rs, int mode, bf16* patches, float* chw, cudaStream_t st) {
    if (n <= 0) return 0;
    if ((H & 1) || (W & 1)) return b200_fail(h, B200CLIP_E_SHAPE, "preprocess_nv12: width and height must be even (%dx%d)", W, H);
    const Plan* pp = nullptr;
    int rc = get_plan(h, H, W, mode, &pp);
    if (rc) return rc;
    const Plan& p = *pp;
    int x0, x1, y0, y1;
    if ((rc = preprocess_source_window_nv12(h, H, W, mode, &x0, &x1, &y0, &y1))) return rc;
    // strides are checked against the window (a compacted upload of it is a valid input)
    if (rs < x1 - x0 || y_fs < rs * (y1 - y0) || uv_fs < rs * (((y1 + 1) >> 1) - (y0 >> 1)))
        return b200_fail(h, B200CLIP_E_ARG, "preprocess_nv12: bad frame geometry %dx%d strides %lld/%lld/%lld", W, H,
                         (long long)rs, (long long)y_fs, (long long)uv_fs);
    if (n > 65535) return b200_fail(h, B200CLIP_E_SHAPE, "preprocess: at most 65535 frames per call (got %d)", n);
    const int S = h->cfg.image_size;
    const int ny = p.ry1 - p.ry0, nx = p.rx1 - p.rx0;
    const int segpx = x1 - x0;
    constexpr int NCV = 160;        // consumers of the fused NV12 kernel: 8 px, two area and two Pillow columns each
    const bool aligned = ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(uv) | static_cast<uintptr_t>(rs) |
                           static_cast<uintptr_t>(y_fs) | static_cast<uintptr_t>(uv_fs)) & 15) == 0 && (segpx & 15) == 0;
    const bool fused = !b200_knobs().nv12_unfused && !b200_knobs().k1_unfused && p.has_a && p.has_b && !p.a_fast && p.a_seq &&
                       p.a_max_cx <= 5 && p.b_max_cnt <= 7 && p.a_int && p.aq != nullptr && aligned && segpx <= 8 * NCV &&
                       nx <= 2 * NCV && S <= 2 * NCV && p.a_dy * 255 < 65536 && p.a_dx <= 255;
    if (fused) {
        const size_t need = p.mid2_per_frame * static_cast<size_t>(n);
        if (need > h->ws_pre_bytes) {
            B200_CUDA(h, cudaStreamSynchronize(st));
            if (h->ws_pre) cudaFree(h->ws_pre);
            h->ws_pre = nullptr; h->ws_pre_bytes = 0;
            B200_CUDA(h, cudaMalloc(&h->ws_pre, need));
            h->ws_pre_bytes = need;
        }
        uint8_t* mid2 = h->ws_pre;
        int rows = 24;
        while (rows > 4 && static_cast<int64_t>(n) * ((ny + rows - 1) / rows) < static_cast<int64_t>(h->num_sms) * 6)
            rows = (rows + 1) / 2;
        const int seg_bytes = 3 * segpx, stage_bytes = 2 * segpx, arow_pitch = (nx * 3 + 32 + 15) & ~15, vpitch = seg_bytes + 32;
        Plan::StripTable stt;
        if ((rc = get_strip_table(h, p, rows, &stt))) return rc;
        const int max_rows = stt.max_rows;
        const size_t smem = 2 * AH_NSTAGE * sizeof(uint64_t) + static_cast<size_t>(AH_NSTAGE) * stage_bytes +
                            4 * static_cast<size_t>(vpitch) + 2 * static_cast<size_t>(arow_pitch) +
                            2 * static_cast<size_t>(max_rows) * sizeof(AhRowInfo);
        if (smem <= 200 * 1024) {
            AhIntParams ip{p.a_dx, p.a_dy, p.a_dx * p.a_dy, p.a_div_shift, p.a_div_mul};
            auto kern = area_hpass_vfirst_kernel<NCV + 32, 3, 1, 2, 2, AH_NSTAGE, true>;
            if (!(h->attr_done & ATTR_K1_NV12)) {
                B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                B200_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                h->attr_done |= ATTR_K1_NV12;
            }
            // algorithmic bytes: the NV12 frame in (1.5 B/px), the patch rows (or fp32 CHW) out
            ProfScope ps(h, PROF_PRE, static_cast<double>(n) * (static_cast<double>(H) * W * 1.5 +
                         (patches ? static_cast<double>(h->grid) * h->grid * h->patch_k * 2.0 : 0.0) +
                         (chw ? 3.0 * S * S * 4.0 : 0.0)), st);
            {
                ProfScope psa(h, PROF_PRE_A, static_cast<double>(n) * (static_cast<double>(y1 - y0) * segpx * 1.5 + ny * S * 3.0), st);
                const int nitems = stt.nstrips * n;
                const int fgrid = persistent_grid(h, kern, NCV + 32, smem, nitems, &h->k1_ctas_per_sm[1]);
                kern<<<fgrid, NCV + 32, smem, st>>>(y, y_fs, rs, uv, uv_fs, segpx, mid2, p.mid2_per_frame, p.ry0, ny, p.rx0, nx, rows,
                                                   x0, seg_bytes, stage_bytes, arow_pitch, vpitch, p.left, S, p.ax, p.ay, p.bx, ip,
                                                   p.aq, stt.nstrips, nitems, max_rows, static_cast<const AhRowInfo*>(stt.rinfo),
                                                   static_cast<const int2*>(stt.meta), 0);
                h->launches++;
            }
            return run_stage_c(h, p, mid2, p.mid2_per_frame, static_cast<int64_t>(S) * 3, 0, p.ry0, n, patches, chw, st);
        }
    }
    // generic form: convert the window to packed RGB in a bounded scratch buffer, group by group, and hand each group
    // to the RGB chain with a virtual frame origin in front of the compacted window
    const int xe = x0 & ~1, pairs = (x1 - xe + 1) >> 1, wrows = y1 - y0;
    const size_t pitch = (static_cast<size_t>(pairs) * 6 + 63) & ~size_t(63), fpitch = pitch * wrows, lead = 64;
    int group = static_cast<int>((size_t(256) << 20) / fpitch);
    if (group < 1) group = 1;
    if (group > n) group = n;
    const size_t need = static_cast<size_t>(group) * fpitch + 2 * lead;
    if (need > h->ws_nv12_bytes) {
        B200_CUDA(h, cudaStreamSynchronize(st));
        if (h->ws_nv12) cudaFree(h->ws_nv12);
        h->ws_nv12 = nullptr; h->ws_nv12_bytes = 0;
        B200_CUDA(h, cudaMalloc(&h->ws_nv12, need));
        h->ws_nv12_bytes = need;
    }
    const size_t prow = static_cast<size_t>(h->grid) * h->grid * h->patch_k;
    for (int i0 = 0; i0 < n; i0 += group) {
        const int nc = (n - i0) < group ? (n - i0) : group;
        uint8_t* stage = h->ws_nv12 + lead;
        {
            ProfScope ps(h, PROF_MISC, static_cast<double>(nc) * wrows * pairs * (3.0 + 6.0), st);
            dim3 grid((static_cast<unsigned>(pairs) * wrows + 255) / 256, nc);
            nv12_window_to_rgb_kernel<<<grid, 256, 0, st>>>(y + i0 * y_fs, uv + i0 * uv_fs, y_fs, uv_fs, rs, stage,
                                                           static_cast<int64_t>(fpitch), static_cast<int64_t>(pitch), xe, y0,
                                                           pairs, wrows);
            h->launches++;
        }
        const uint8_t* origin = stage - static_cast<int64_t>(y0) * static_cast<int64_t>(pitch) - static_cast<int64_t>(xe) * 3;
        rc = launch_preprocess(h, origin, nc, H, W, static_cast<int64_t>(fpitch), static_cast<int64_t>(pitch), mode,
                               patches ? patches + static_cast<size_t>(i0) * prow : nullptr,
                               chw ? chw + static_cast<size_t>(i0) * 3 * S * S : nullptr, st);
        if (rc) return rc;
    }
    B200_CUDA(h, cudaGetLastError());
    return 0;
}
