// spectral.cu -- element-wise parts of the FFT derivatives (the transforms themselves are issued by solver.cu).
//
// Replaces (reference, L/ = CuSuperHelium/CuSuperHelium/):
//   batched_vector_subtract_singletime_complex_real   L/utilities.cuh:231-239   (launched L/Derivatives.cuh:317-318)
//   first_derivative_multiplication                   L/utilities.cuh:106-148
//   second_derivative_fft                             L/utilities.cuh:178-200
//   vector_mutiply_scalar / vector_scalar_add_complex_real   L/utilities.cuh:249-263 (L/Derivatives.cuh:254, 374, 382)
// The Nyquist conventions of the reference are reproduced exactly:
//   D1: modes 0..n/2-1 -> i k; mode n/2 -> i * pi * (n/2); mode n/2+1 -> 0; modes above -> i (k - n)
//   D2: -(k^2) for every mode, nothing zeroed.  The 1/n of the unnormalised inverse is folded into the multiply.
#include "internal.cuh"

namespace rb {

__device__ __forceinline__ double2 d1_coeff(double2 c, int i, int n) {
    const double dn = static_cast<double>(n);
    if (i < n / 2) {
        return make_double2(-i * c.y / dn, i * c.x / dn);
    } else if (i == n / 2) {
        return make_double2(-kPi * i * c.y / dn, kPi * i * c.x / dn);
    } else if (i == n / 2 + 1) {
        return make_double2(0.0, 0.0);
    }
    return make_double2(-(i - n) * c.y / dn, (i - n) * c.x / dn);
}

__device__ __forceinline__ double2 d2_coeff(double2 c, int i, int n) {
    if (i <= n / 2) {
        return make_double2(-i * i * c.x / (n), -i * i * c.y / (n));
    }
    return make_double2(-(i - n) * (i - n) * c.x / (n), -(i - n) * (i - n) * c.y / (n));
}

// Z - 2 pi j / N  and  Phi - (-(1+rho) pi U / N) j      (ZPhiDerivative ctor tables, L/Derivatives.cuh:282-287)
__global__ void sub_linear_kernel(const double2* __restrict__ Z, const double2* __restrict__ Phi, double2* __restrict__ zper,
                                  double2* __restrict__ phiper, int N, size_t total, double rho, double U) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int i = (int)(tid % N);
    double zl = 2 * kPi * (double)i / N;
    double2 z = Z[tid];
    zper[tid] = make_double2(z.x - zl, z.y);
    if (Phi) {
        double pl = -(1 + rho) * kPi * U / N * (double)i;
        double2 p = Phi[tid];
        phiper[tid] = make_double2(p.x - pl, p.y);
    }
}

void launch_sub_linear(const double2* Z, const double2* Phi, double2* out_zper, double2* out_phiper, int N, int batch,
                       double rho, double U, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    sub_linear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Z, Phi, out_zper, out_phiper, N, total, rho, U);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// one pass over the spectra of Z_per and Phi_per: i k Z^, -k^2 Z^, i k Phi^
__global__ void spectral_multiply_zphi_kernel(const double2* __restrict__ hatZ, const double2* __restrict__ hatPhi,
                                              double2* __restrict__ d1z, double2* __restrict__ d2z, double2* __restrict__ d1phi,
                                              int N, size_t total) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int i = (int)(tid % N);
    double2 cz = hatZ[tid];
    d1z[tid] = d1_coeff(cz, i, N);
    d2z[tid] = d2_coeff(cz, i, N);
    if (hatPhi) d1phi[tid] = d1_coeff(hatPhi[tid], i, N);
}

void launch_spectral_multiply_zphi(const double2* hatZ, const double2* hatPhi, double2* out_d1z, double2* out_d2z,
                                   double2* out_d1phi, int N, int batch, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    spectral_multiply_zphi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(hatZ, hatPhi, out_d1z, out_d2z, out_d1phi, N,
                                                                                   total);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void spectral_multiply_kernel(const double2* __restrict__ hat, double2* __restrict__ out, int N, size_t total,
                                         int second) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    int i = (int)(tid % N);
    double2 c = hat[tid];
    out[tid] = second ? d2_coeff(c, i, N) : d1_coeff(c, i, N);
}

void launch_spectral_multiply(const double2* hat, double2* out, int N, int batch, int second, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    spectral_multiply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(hat, out, N, total, second);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// first derivative of a REAL sequence from its half spectrum (D2Z transform): the full coefficient array of the reference's
// Z2Z path is rebuilt with c[N-i] = conj(c[i]); `scale` (2 pi / N) is folded into the coefficients
__global__ void spectral_multiply_real_kernel(const double2* __restrict__ half, double2* __restrict__ out, int N, size_t total,
                                              double scale) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const int nh = N / 2 + 1;
    size_t b = tid / N;
    int i = (int)(tid - b * N);
    double2 c;
    if (i < nh) {
        c = half[b * nh + i];
    } else {
        c = half[b * nh + (N - i)];
        c.y = -c.y;
    }
    double2 r = d1_coeff(c, i, N);
    out[tid] = make_double2(r.x * scale, r.y * scale);
}

void launch_spectral_multiply_real(const double2* half, double2* out, int N, int batch, double scale, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    spectral_multiply_real_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(half, out, N, total, scale);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// Shared-memory FFT derivatives for power-of-two N <= 4096 (one CTA per transform pair): forward transform, coefficient
// multiply and inverse transform in ONE launch.  At these sizes the library transforms are launch-bound (ncu, N = 4096:
// 6.5-7.8 us per cuFFT launch, 6 launches per RHS); here the whole derivative costs about one such launch.
// Radix-2 Stockham autosort (no bit reversal), twiddles exp(-2 pi i k / N) from a host-computed table staged in shared memory.
// Unnormalised in both directions, like cuFFT: the 1/n sits in the coefficient multiply.
// ------------------------------------------------------------------------------------------------
// stage the twiddle table in shared memory (the passes are latency-bound if every pass goes to L2 for its twiddles)
// table layout: for every pass Ns = 1, 2, 4, .., N/2 the Ns twiddles exp(-i pi k / Ns), k < Ns, stored contiguously at offset
// Ns - 1 (N - 1 entries in all), so that a warp reads consecutive entries (a strided walk through one exp(-2 pi i k / N) table
// is a 32-way bank conflict for most passes)
__device__ __forceinline__ void stage_twiddles(double2* sm_tw, const double2* __restrict__ tw, int N) {
    for (int i = threadIdx.x; i < N - 1; i += blockDim.x) sm_tw[i] = tw[i];
}

// Stockham autosort radix-2: natural order in, natural order out, ping-pong between two shared buffers, no bit reversal
// (a bit-reversed scatter in shared memory is a 32-way bank conflict).  Returns the buffer that holds the result.
__device__ __forceinline__ double2* fft_stockham(double2* __restrict__ a, double2* __restrict__ b, int N,
                                                 const double2* __restrict__ tw, bool inverse) {
    const int T = blockDim.x;
    const int halfN = N >> 1;
    for (int Ns = 1; Ns < N; Ns <<= 1) {
        const double2* __restrict__ twp = tw + (Ns - 1);
        for (int j = threadIdx.x; j < halfN; j += T) {
            const int k = j & (Ns - 1);
            double2 w = twp[k];
            if (inverse) w.y = -w.y;
            const double2 v0 = a[j];
            const double2 v1 = a[j + halfN];
            const double2 t = make_double2(v1.x * w.x - v1.y * w.y, v1.x * w.y + v1.y * w.x);
            const int j0 = ((j - k) << 1) + k;
            b[j0] = make_double2(v0.x + t.x, v0.y + t.y);
            b[j0 + Ns] = make_double2(v0.x - t.x, v0.y - t.y);
        }
        __syncthreads();
        double2* tmp = a;
        a = b;
        b = tmp;
    }
    return a;
}

// blockIdx.x: 0 -> i k Z^ (Zp), 1 -> -k^2 Z^ (Zpp), 2 -> i k Phi^ (PhiPrime); blockIdx.y: batch member.
// Outputs are the raw inverse transforms (scaling and linear parts are applied by the geometry / finish kernel).
__global__ void fft_zphi_kernel(const double2* __restrict__ Z, const double2* __restrict__ Phi, double2* __restrict__ Zp,
                                double2* __restrict__ Zpp, double2* __restrict__ PhiP, int N, int logN,
                                const double2* __restrict__ tw, double rho, double U) {
    extern __shared__ double2 sm_fft[];
    double2* bufA = sm_fft;
    double2* bufB = sm_fft + N;
    double2* stw = sm_fft + 2 * N;
    const int role = blockIdx.x;
    const size_t off = (size_t)blockIdx.y * N;
    const double2* in = (role == 2 ? Phi : Z) + off;
    double2* out = (role == 0 ? Zp : (role == 1 ? Zpp : PhiP)) + off;
    stage_twiddles(stw, tw, N);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double2 v = in[i];
        const double lin = role == 2 ? -(1 + rho) * kPi * U / N * (double)i : 2 * kPi * (double)i / N;
        v.x -= lin;
        bufA[i] = v;
    }
    __syncthreads();
    double2* r = fft_stockham(bufA, bufB, N, stw, false);
    double2* o = (r == bufA) ? bufB : bufA;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double2 c = r[i];
        r[i] = role == 1 ? d2_coeff(c, i, N) : d1_coeff(c, i, N);
    }
    __syncthreads();
    r = fft_stockham(r, o, N, stw, true);
    for (int i = threadIdx.x; i < N; i += blockDim.x) out[i] = r[i];
}

// a' = scale * D1(x) for a real vector x (one CTA per batch member); returns at once when the solve it belongs to is finished
__global__ void fft_real_derivative_kernel(const double* __restrict__ x, double2* __restrict__ out, int N, int logN,
                                           const double2* __restrict__ tw, double scale, const SolveCtrl* ctrl) {
    extern __shared__ double2 sm_fft[];
    if (ctrl && *reinterpret_cast<const volatile int*>(&ctrl->done)) return;
    double2* bufA = sm_fft;
    double2* bufB = sm_fft + N;
    double2* stw = sm_fft + 2 * N;
    const size_t off = (size_t)blockIdx.y * N;
    stage_twiddles(stw, tw, N);
    for (int i = threadIdx.x; i < N; i += blockDim.x) bufA[i] = make_double2(x[off + i], 0.0);
    __syncthreads();
    double2* r = fft_stockham(bufA, bufB, N, stw, false);
    double2* o = (r == bufA) ? bufB : bufA;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double2 c = d1_coeff(r[i], i, N);
        r[i] = make_double2(c.x * scale, c.y * scale);
    }
    __syncthreads();
    r = fft_stockham(r, o, N, stw, true);
    for (int i = threadIdx.x; i < N; i += blockDim.x) out[off + i] = r[i];
}

// ------------------------------------------------------------------------------------------------
// Radix-8 register-resident transform for 256 <= N <= 8192 (N a power of two): N/8 threads, every thread owns the 8 elements at
// positions j + r N/8 for the WHOLE derivative (forward transform, coefficient multiply, inverse transform).  Stockham autosort
// with passes of radix 2^(log2 N mod 3) (first, trivial twiddles, done on the registers the thread already holds) and radix 8
// (ceil-free: 4 passes at N = 4096 against 12 radix-2 passes); between passes the results are exchanged through ONE padded
// shared-memory array (position p lives at p + p/8: the stride-8 scatter of the early passes is conflict-free), two barriers per
// exchange.  The last pass leaves the natural-order result in exactly the register layout the first pass of the next transform
// reads, and global memory is read and written straight from / to registers (coalesced: consecutive threads, consecutive elements).
// Twiddles: w and w^2 of a butterfly are table entries (2 Ns per pass staged in shared memory), the other five products.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiplication by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ double2 mul_mi(double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }

template <bool INV>
__device__ __forceinline__ void dft4(double2& a0, double2& a1, double2& a2, double2& a3) {
    const double2 s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a1 = cadd(d02, d13);
    a2 = csub(s02, s13);
    a3 = csub(d02, d13);
}

template <bool INV>
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
    dft4<INV>(v[0], v[2], v[4], v[6]);   // even part -> E0..E3 in v[0], v[2], v[4], v[6]
    dft4<INV>(v[1], v[3], v[5], v[7]);   // odd part  -> O0..O3 in v[1], v[3], v[5], v[7]
    constexpr double h = 0.70710678118654752440;
    // W8^k O_k, W8 = exp(-+ 2 pi i / 8)
    const double2 o0 = v[1];
    const double2 o1 = INV ? make_double2(h * (v[3].x - v[3].y), h * (v[3].x + v[3].y)) : make_double2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));
    const double2 o2 = mul_mi<INV>(v[5]);
    const double2 o3 = INV ? make_double2(-h * (v[7].x + v[7].y), h * (v[7].x - v[7].y)) : make_double2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));
    const double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    v[0] = cadd(e0, o0);
    v[1] = cadd(e1, o1);
    v[2] = cadd(e2, o2);
    v[3] = cadd(e3, o3);
    v[4] = csub(e0, o0);
    v[5] = csub(e1, o1);
    v[6] = csub(e2, o2);
    v[7] = csub(e3, o3);
}

__device__ __forceinline__ int fft8_pad(int p) { return p + (p >> 3); }

// number of staged twiddles: 2 Ns per radix-8 pass (w and w^2 of every butterfly are table entries)
__host__ __device__ inline int fft8_twiddle_count(int N, int logN) {
    int n = 0;
    for (int Ns = 1 << (logN % 3); Ns * 8 <= N; Ns <<= 3) n += 2 * Ns;
    return n;
}

// tw: the per-pass table of solver_create (exp(-i pi k / Ns') at offset Ns' - 1); the radix-8 pass at Ns uses Ns' = 4 Ns:
// exp(-2 pi i m / (8 Ns)), of which m < 2 Ns are staged.  The passes' pieces are loaded as ONE batch per thread (all loads in
// flight together with the caller's input loads: one L2 round trip), then stored.
constexpr int kFft8TwBatch = 4;
__device__ __forceinline__ void fft8_stage_twiddles(double2* stw, const double2* __restrict__ tw, int N, int logN) {
    const int total = fft8_twiddle_count(N, logN);
    for (int i0 = threadIdx.x; i0 < total; i0 += kFft8TwBatch * blockDim.x) {
        double2 w[kFft8TwBatch];
#pragma unroll
        for (int u = 0; u < kFft8TwBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            w[u] = make_double2(0.0, 0.0);
            if (i < total) {
                int off = 0, Ns = 1 << (logN % 3);
                while (i >= off + 2 * Ns) {   // which pass this entry belongs to
                    off += 2 * Ns;
                    Ns <<= 3;
                }
                w[u] = tw[4 * Ns - 1 + (i - off)];
            }
        }
#pragma unroll
        for (int u = 0; u < kFft8TwBatch; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < total) stw[i] = w[u];
        }
    }
}

// v[r] = element j + r N/8 on entry and on return (natural order); sm: one (N <= 4096: two, ping-pong) padded arrays of N + N/8
// entries; the caller has synchronised stw.  Twiddles of a butterfly: w = exp(-+ 2 pi i k / (8 Ns)) and w^2 from the table, the
// other five by one or two multiplications (w^3 = w w^2, w^4 = (w^2)^2, w^5 = w^4 w, w^6 = w^4 w^2, w^7 = w^4 w^3) -- seven table
// reads per butterfly cost more shared-memory wavefronts than the whole exchange (even strides conflict), and the accuracy is the
// same to the last digit the tests can see.
template <bool INV>
__device__ __forceinline__ void fft8_transform(double2 (&v)[8], double2* __restrict__ sm, const double2* __restrict__ stw, int N, int logN,
                                               bool two_buffers) {
    const int j = threadIdx.x, T = N >> 3;
    const int b = logN % 3;
    double2* cur = sm;
    double2* const other = two_buffers ? sm + N + (N >> 3) : sm;
    auto exchange = [&]() {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = cur[fft8_pad(j + r * T)];
        if (!two_buffers) __syncthreads();      // (ping-pong: the next writes go to the other array, no second barrier)
        cur = (cur == sm) ? other : sm;
    };
    if (b == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int jj = j + q * T;
            cur[fft8_pad(2 * jj)] = cadd(v[q], v[q + 4]);
            cur[fft8_pad(2 * jj + 1)] = csub(v[q], v[q + 4]);
        }
        exchange();
    } else if (b == 2) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int jj = j + q * T;
            dft4<INV>(v[q], v[q + 2], v[q + 4], v[q + 6]);
#pragma unroll
            for (int m = 0; m < 4; ++m) cur[fft8_pad(4 * jj + m)] = v[q + 2 * m];
        }
        exchange();
    }
    int off = 0;
    for (int Ns = 1 << b; Ns * 8 <= N; Ns <<= 3) {
        const int k = j & (Ns - 1);
        if (Ns > 1) {
            double2 w1 = stw[off + k], w2 = stw[off + 2 * k];
            if (INV) {
                w1.y = -w1.y;
                w2.y = -w2.y;
            }
            const double2 w3 = cmul(w1, w2), w4 = cmul(w2, w2);
            v[1] = cmul(v[1], w1);
            v[2] = cmul(v[2], w2);
            v[3] = cmul(v[3], w3);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w1));
            v[6] = cmul(v[6], cmul(w4, w2));
            v[7] = cmul(v[7], cmul(w4, w3));
        }
        dft8<INV>(v);
        if (Ns * 8 < N) {
            const int j0 = ((j - k) << 3) + k;
#pragma unroll
            for (int r = 0; r < 8; ++r) cur[fft8_pad(j0 + r * Ns)] = v[r];
            exchange();
        }
        off += 2 * Ns;
    }
    // (an even number of exchanges is not guaranteed: the next transform simply starts on whichever array is current -- both are
    // free here, every thread has passed the last barrier after its last read only if a barrier follows; see the callers)
}

// d1_coeff / d2_coeff for power-of-two n: 1/n is exact, so the multiplication gives the quotient bit for bit (an FP64 division is
// ~30 instructions; 24 of them per thread were a fifth of the kernel)
__device__ __forceinline__ double2 d1_coeff_pow2(double2 c, int i, int n, double inv_n) {
    if (i < n / 2) return make_double2((-i * c.y) * inv_n, (i * c.x) * inv_n);
    if (i == n / 2) return make_double2((-kPi * i * c.y) * inv_n, (kPi * i * c.x) * inv_n);
    if (i == n / 2 + 1) return make_double2(0.0, 0.0);
    return make_double2((-(i - n) * c.y) * inv_n, ((i - n) * c.x) * inv_n);
}

__device__ __forceinline__ double2 d2_coeff_pow2(double2 c, int i, int n, double inv_n) {
    const int k = i <= n / 2 ? i : i - n;
    return make_double2((-k * k * c.x) * inv_n, (-k * k * c.y) * inv_n);
}

// the three derivatives of one surface, as fft_zphi_kernel: blockIdx.x = role, blockIdx.y = batch member; N/8 threads
template <int MAXT>
__global__ void __launch_bounds__(MAXT) fft8_zphi_kernel(const double2* __restrict__ Z, const double2* __restrict__ Phi, double2* __restrict__ Zp,
                                                         double2* __restrict__ Zpp, double2* __restrict__ PhiP, int N, int logN,
                                                         const double2* __restrict__ tw, double rho, double U) {
    extern __shared__ double2 sm_fft[];
    const bool two = N <= 4096;
    double2* stw = sm_fft + (two ? 2 : 1) * (N + (N >> 3));
    const int role = blockIdx.x;
    const size_t off = (size_t)blockIdx.y * N;
    const double2* in = (role == 2 ? Phi : Z) + off;
    double2* out = (role == 0 ? Zp : (role == 1 ? Zpp : PhiP)) + off;
    const int j = threadIdx.x, T = N >> 3;
    const double inv_n = 1.0 / (double)N;
    double2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = in[j + r * T];
    fft8_stage_twiddles(stw, tw, N, logN);     // (its loads fly together with the input loads above)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = j + r * T;
        v[r].x -= role == 2 ? -(1 + rho) * kPi * U / N * (double)i : (2 * kPi * (double)i) * inv_n;   // (= ... / N: N is a power of two)
    }
    __syncthreads();
    fft8_transform<false>(v, sm_fft, stw, N, logN, two);
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = role == 1 ? d2_coeff_pow2(v[r], j + r * T, N, inv_n) : d1_coeff_pow2(v[r], j + r * T, N, inv_n);
    __syncthreads();                           // every read of the forward transform's last exchange is done
    fft8_transform<true>(v, sm_fft, stw, N, logN, two);
#pragma unroll
    for (int r = 0; r < 8; ++r) out[j + r * T] = v[r];
}

// a' = scale * D1(x) for a real vector x, as fft_real_derivative_kernel
template <int MAXT>
__global__ void __launch_bounds__(MAXT) fft8_real_derivative_kernel(const double* __restrict__ x, double2* __restrict__ out, int N, int logN,
                                                                    const double2* __restrict__ tw, double scale, const SolveCtrl* ctrl) {
    extern __shared__ double2 sm_fft[];
    if (ctrl && *reinterpret_cast<const volatile int*>(&ctrl->done)) return;
    const bool two = N <= 4096;
    double2* stw = sm_fft + (two ? 2 : 1) * (N + (N >> 3));
    const size_t off = (size_t)blockIdx.y * N;
    const int j = threadIdx.x, T = N >> 3;
    double2 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = make_double2(x[off + j + r * T], 0.0);
    fft8_stage_twiddles(stw, tw, N, logN);
    __syncthreads();
    fft8_transform<false>(v, sm_fft, stw, N, logN, two);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const double2 c = d1_coeff_pow2(v[r], j + r * T, N, 1.0 / (double)N);
        v[r] = make_double2(c.x * scale, c.y * scale);
    }
    __syncthreads();
    fft8_transform<true>(v, sm_fft, stw, N, logN, two);
#pragma unroll
    for (int r = 0; r < 8; ++r) out[off + j + r * T] = v[r];
}

static bool fft8_ok(int N) { return N >= 256 && N <= 8192; }
static size_t fft8_smem(int N, int logN) { return (size_t)((N <= 4096 ? 2 : 1) * (N + (N >> 3)) + fft8_twiddle_count(N, logN)) * sizeof(double2); }

static int fft_threads(int N) { return N / 2 >= 1024 ? 1024 : (N / 2 >= 32 ? N / 2 : 32); }

static void fft_smem_attr(const void* fn, size_t bytes) {
    if (bytes > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void launch_fft_zphi(const double2* Z, const double2* Phi, double2* Zp, double2* Zpp, double2* PhiP, int N, int logN, int batch,
                     const double2* tw, double rho, double U, cudaStream_t st) {
    if (fft8_ok(N)) {
        const size_t b8 = fft8_smem(N, logN);
        auto kern = N <= 4096 ? fft8_zphi_kernel<512> : fft8_zphi_kernel<1024>;   // 512 threads: no register cap, no spills
        fft_smem_attr((const void*)kern, b8);
        kern<<<dim3(Phi ? 3 : 2, batch), N / 8, b8, st>>>(Z, Phi, Zp, Zpp, PhiP, N, logN, tw, rho, U);
        RB_CUDA(cudaGetLastError());
        count_launch();
        return;
    }
    const size_t bytes = (size_t)(3 * N) * sizeof(double2);
    fft_smem_attr((const void*)fft_zphi_kernel, bytes);
    fft_zphi_kernel<<<dim3(Phi ? 3 : 2, batch), fft_threads(N), bytes, st>>>(Z, Phi, Zp, Zpp, PhiP, N, logN, tw, rho, U);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

void launch_fft_real_derivative(const double* x, double2* out, int N, int logN, int batch, const double2* tw, double scale,
                                const SolveCtrl* ctrl, cudaStream_t st) {
    if (fft8_ok(N)) {
        const size_t b8 = fft8_smem(N, logN);
        auto kern = N <= 4096 ? fft8_real_derivative_kernel<512> : fft8_real_derivative_kernel<1024>;
        fft_smem_attr((const void*)kern, b8);
        kern<<<dim3(1, batch), N / 8, b8, st>>>(x, out, N, logN, tw, scale, ctrl);
        RB_CUDA(cudaGetLastError());
        count_launch();
        return;
    }
    const size_t bytes = (size_t)(3 * N) * sizeof(double2);
    fft_smem_attr((const void*)fft_real_derivative_kernel, bytes);
    fft_real_derivative_kernel<<<dim3(1, batch), fft_threads(N), bytes, st>>>(x, out, N, logN, tw, scale, ctrl);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

// scaling by 2 pi / N (resp. its square) and the linear parts put back (L/Derivatives.cuh:321-324, 374, 380-383)
__global__ void finish_zphi_kernel(double2* __restrict__ Zp, double2* __restrict__ Zpp, double2* __restrict__ PhiP, int N,
                                   size_t total, double rho, double U) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= total) return;
    const double s1 = 2.0 * kPi / N;
    const double s2 = 4.0 * kPi * kPi / ((double)N * N);
    double2 a = Zp[tid];
    Zp[tid] = make_double2(a.x * s1 + 2.0 * kPi / N, a.y * s1);
    double2 c = Zpp[tid];
    Zpp[tid] = make_double2(c.x * s2, c.y * s2);
    if (PhiP) {
        double2 p = PhiP[tid];
        p.x *= s1;
        p.y *= s1;
        if (U != 0) p.x += -(1 + rho) * kPi * U / N;
        PhiP[tid] = p;
    }
}

void launch_finish_zphi(double2* Zp, double2* Zpp, double2* PhiP, int N, int batch, double rho, double U, cudaStream_t st) {
    size_t total = (size_t)N * batch;
    finish_zphi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Zp, Zpp, PhiP, N, total, rho, U);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

__global__ void scale_kernel(double2* __restrict__ v, double s, size_t n) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    double2 a = v[tid];
    v[tid] = make_double2(a.x * s, a.y * s);
}

void launch_scale(double2* v, double s, size_t n, cudaStream_t st) {
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, s, n);
    RB_CUDA(cudaGetLastError());
    count_launch();
}

}  // namespace rb
