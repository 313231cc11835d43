// Observables on the hot path: unfolded samples in id order, block MSD, density modes, thermodynamic sums.
//
// Reference: halmd/observables/gpu/phase_space_kernel.cu:39-66; halmd/observables/gpu/dynamics/
// mean_square_displacement.cpp:55-68 with tagged_particle.hpp:33-61 and halmd/numeric/accumulator.hpp:56-132;
// halmd/observables/gpu/density_mode_kernel.cu:46-108; halmd/observables/gpu/thermodynamics_kernel.cu:35-97.
//
// The reference launches one 32x256 reduction (+ blocking D2H) per lag time; here one launch correlates the
// first sample of a block with all `nlags` later samples and leaves (count, mean, M2) triples in device memory.
#include "common.cuh"

namespace hx {

__global__ void k_sample_positions(const float4* __restrict__ pos, const float4* __restrict__ image,
                                   const unsigned* __restrict__ rid, unsigned n, vec3 L, float4* __restrict__ out)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned r = rid[i];
    float4 p = __ldg(&pos[r]);
    float4 im = __ldg(&image[r]);
    // extend_periodic: r += image * L, contracted to FMA by nvcc in the reference (box_kernel.cuh:54-61)
    out[i] = make_float4(__fmaf_rn(im.x, L.x, p.x), __fmaf_rn(im.y, L.y, p.y), __fmaf_rn(im.z, L.z, p.z), p.w);
}

__global__ void k_sample_velocities(const float4* __restrict__ vel, const unsigned* __restrict__ rid, unsigned n,
                                    float4* __restrict__ out)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __ldg(&vel[rid[i]]);
}

// Welford / Chan accumulators in double
struct wf { double n, m, v; };

__device__ __forceinline__ void wf_add(wf& a, double x)
{
    double t = x - a.m;
    a.n += 1.0;
    a.m += t / a.n;
    a.v += t * (x - a.m);
}

__device__ __forceinline__ void wf_merge(wf& a, wf const& b)
{
    // accumulator.hpp:104-118
    if (b.n > 0) {
        double count = a.n + b.n;
        double diff = a.m - b.m;
        a.v += b.v + diff * diff * a.n * b.n / count;
        a.m = (a.n * a.m + b.n * b.m) / count;
        a.n = count;
    }
}

__device__ __forceinline__ wf wf_shfl_xor(wf a, int o)
{
    wf b;
    b.n = __shfl_xor_sync(0xFFFFFFFFu, a.n, o);
    b.m = __shfl_xor_sync(0xFFFFFFFFu, a.m, o);
    b.v = __shfl_xor_sync(0xFFFFFFFFu, a.v, o);
    return b;
}

// inner_prod(a, b) in the order nvcc emits for the reference (cf. norm2 in common.cuh): fma(z, z', fma(x, x', y * y'))
__device__ __forceinline__ float norm2_mixed(float4 a, float4 b)
{
    return __fmaf_rn(a.z, b.z, __fmaf_rn(a.x, b.x, __fmul_rn(a.y, b.y)));
}

constexpr unsigned MSD_BLOCKS = 128;
constexpr unsigned MSD_THREADS = 256;

// correlation functions of two tagged-particle samples (halmd/observables/dynamics/*.hpp), evaluated in float
//   HX_CORR_MSD   |r2 - r1|^2        mean_square_displacement.hpp:35-41
//   HX_CORR_MQD   (|r2 - r1|^2)^2    mean_quartic_displacement.hpp:36-44
//   HX_CORR_VACF  v1 . v2            velocity_autocorrelation.hpp:36-39
template <int KIND>
__device__ __forceinline__ float correlate(float4 a, float4 b)
{
    if (KIND == HX_CORR_VACF) return norm2_mixed(a, b);
    float dx = __fsub_rn(b.x, a.x), dy = __fsub_rn(b.y, a.y), dz = __fsub_rn(b.z, a.z);
    float rr = norm2(dx, dy, dz);
    return KIND == HX_CORR_MQD ? __fmul_rn(rr, rr) : rr;
}

// grid = (MSD_BLOCKS, nlags): block accumulators → partial[(lag * MSD_BLOCKS + block)]
template <int KIND>
__global__ void k_msd_partial(const float4* __restrict__ first, const float4* const* __restrict__ samples, unsigned n,
                              wf* __restrict__ partial)
{
    __shared__ wf s_w[MSD_THREADS / 32];
    const float4* second = samples[blockIdx.y];
    wf acc = {0., 0., 0.};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        wf_add(acc, static_cast<double>(correlate<KIND>(first[i], second[i])));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wf b = wf_shfl_xor(acc, o);
        // merge in a lane-symmetric way so that all lanes end with the same value
        wf lo = (threadIdx.x & o) ? b : acc, hi = (threadIdx.x & o) ? acc : b;
        wf_merge(lo, hi);
        acc = lo;
    }
    if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        wf t = s_w[0];
        for (unsigned w = 1; w < MSD_THREADS / 32; ++w) wf_merge(t, s_w[w]);
        partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void k_msd_final(const wf* __restrict__ partial, unsigned nblocks, double* __restrict__ out)
{
    unsigned lag = blockIdx.x;
    if (threadIdx.x != 0) return;
    wf t = partial[lag * nblocks];
    for (unsigned b = 1; b < nblocks; ++b) wf_merge(t, partial[lag * nblocks + b]);
    out[3 * lag + 0] = t.n;
    out[3 * lag + 1] = t.m;
    out[3 * lag + 2] = t.v;
}

// density modes: block (x) strides over particles, loops over a tile of wavevectors held in registers
constexpr unsigned DM_THREADS = 256;
constexpr unsigned DM_QTILE = 8;

__global__ void k_density_modes(const float4* __restrict__ pos, unsigned n, const float* __restrict__ q, unsigned nq,
                                double* __restrict__ rho)
{
    __shared__ double s_c[DM_THREADS / 32][DM_QTILE], s_s[DM_THREADS / 32][DM_QTILE];
    unsigned const q0 = blockIdx.y * DM_QTILE;
    float qx[DM_QTILE], qy[DM_QTILE], qz[DM_QTILE];
    double c[DM_QTILE], s[DM_QTILE];
#pragma unroll
    for (unsigned k = 0; k < DM_QTILE; ++k) {
        bool ok = q0 + k < nq;
        qx[k] = ok ? q[3 * (q0 + k) + 0] : 0.f;
        qy[k] = ok ? q[3 * (q0 + k) + 1] : 0.f;
        qz[k] = ok ? q[3 * (q0 + k) + 2] : 0.f;
        c[k] = 0.;
        s[k] = 0.;
    }
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pos[i];
#pragma unroll
        for (unsigned k = 0; k < DM_QTILE; ++k) {
            // inner_prod(q, r) in the reference's compiled order, then accurate sincos
            float qr = __fmaf_rn(qz[k], p.z, __fmaf_rn(qx[k], p.x, __fmul_rn(qy[k], p.y)));
            float sn, cs;
            sincosf(qr, &sn, &cs);
            c[k] += cs;
            s[k] += sn;
        }
    }
#pragma unroll
    for (unsigned k = 0; k < DM_QTILE; ++k) {
        c[k] = warp_sum(c[k]);
        s[k] = warp_sum(s[k]);
    }
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (unsigned k = 0; k < DM_QTILE; ++k) {
            s_c[threadIdx.x >> 5][k] = c[k];
            s_s[threadIdx.x >> 5][k] = s[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < DM_QTILE && q0 + threadIdx.x < nq) {
        double cc = 0., ss = 0.;
        for (unsigned w = 0; w < DM_THREADS / 32; ++w) {
            cc += s_c[w][threadIdx.x];
            ss += s_s[w][threadIdx.x];
        }
        // (sum cos, -sum sin): density_mode.cpp:126-129.  Block sums are merged as 31.32 fixed-point integers: integer
        // addition is associative, so the result does not depend on the order in which blocks finish (|sum| <= n < 2^31;
        // the resolution 2^-32 is far below the accuracy of a float sine); k_density_modes_finalise converts in place.
        unsigned long long* const acc = reinterpret_cast<unsigned long long*>(rho);
        atomicAdd(&acc[2 * (q0 + threadIdx.x) + 0], static_cast<unsigned long long>(__double2ll_rn(cc * 4294967296.0)));
        atomicAdd(&acc[2 * (q0 + threadIdx.x) + 1], static_cast<unsigned long long>(__double2ll_rn(-ss * 4294967296.0)));
    }
}

__global__ void k_density_modes_finalise(double* __restrict__ rho, unsigned nwords)
{
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) rho[i] = static_cast<double>(reinterpret_cast<long long*>(rho)[i]) * (1.0 / 4294967296.0);
}

constexpr unsigned TH_THREADS = 256;
constexpr unsigned TH_BLOCKS = HX_THERMO_BLOCKS;

// Thermodynamic sums with the per-particle terms of the reference's accumulators (thermodynamics_kernel.cu:35-97) as its
// build evaluates them: inner_prod(v, v) in float (FMUL + 2 FFMA), the product with the mass fused into the
// double-single accumulation, i.e. exact — here exact in double.  Block partials go to scratch and are combined by the
// last block to finish, in block order: the sums do not depend on the order in which blocks run.
__global__ void k_thermodynamics(const float4* __restrict__ vel, const float* __restrict__ en_pot,
                                 const float* __restrict__ stress, size_t stress_stride, unsigned n,
                                 double* __restrict__ out)
{
    __shared__ double s_red[TH_THREADS / 32][7];
    __shared__ bool s_last;
    double acc[7] = {0., 0., 0., 0., 0., 0., 0.};
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v = vel[i];
        double const m = static_cast<double>(v.w);
        acc[0] += m * static_cast<double>(norm2(v.x, v.y, v.z));
        acc[1] += m * static_cast<double>(v.x);
        acc[2] += m * static_cast<double>(v.y);
        acc[3] += m * static_cast<double>(v.z);
        acc[4] += m;
        if (en_pot) acc[5] += static_cast<double>(en_pot[i]);
        if (stress) {
            acc[6] += static_cast<double>(stress[i]) + static_cast<double>(stress[stress_stride + i])
                      + static_cast<double>(stress[2 * stress_stride + i]);
        }
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) acc[k] = warp_sum(acc[k]);
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) s_red[threadIdx.x >> 5][k] = acc[k];
    }
    __syncthreads();
    double* const partial = out + 8;                  // [TH_BLOCKS][8]
    unsigned* const ticket = reinterpret_cast<unsigned*>(out + 7);
    if (threadIdx.x < 7) {
        double t = 0.;
        for (unsigned w = 0; w < TH_THREADS / 32; ++w) t += s_red[w][threadIdx.x];
        partial[blockIdx.x * 8 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x < 7) {
        __threadfence();
        double t = 0.;
        for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(&partial[b * 8 + threadIdx.x]);
        out[threadIdx.x] = t;
    }
}

} // namespace hx

using namespace hx;

extern "C" {

int hx_sample_positions(const hx_float4* position_hi_dev, const hx_float4* image_dev, const uint32_t* reverse_id_dev,
                        size_t nparticle, const float box_length[3], hx_float4* sample_dev, void* stream)
{
    if (!position_hi_dev || !image_dev || !reverse_id_dev || !box_length || !sample_dev || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_sample_positions: invalid argument");
        return HX_E_INVALID;
    }
    if (nparticle == 0) return HX_OK;
    k_sample_positions<<<div_up(nparticle, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(position_hi_dev), reinterpret_cast<const float4*>(image_dev), reverse_id_dev,
        static_cast<unsigned>(nparticle), make_vec3(box_length), reinterpret_cast<float4*>(sample_dev));
    return check_launch("k_sample_positions");
}

int hx_sample_velocities(const hx_float4* velocity_hi_dev, const uint32_t* reverse_id_dev, size_t nparticle,
                         hx_float4* sample_dev, void* stream)
{
    if (!velocity_hi_dev || !reverse_id_dev || !sample_dev || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_sample_velocities: invalid argument");
        return HX_E_INVALID;
    }
    if (nparticle == 0) return HX_OK;
    k_sample_velocities<<<div_up(nparticle, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(velocity_hi_dev), reverse_id_dev, static_cast<unsigned>(nparticle),
        reinterpret_cast<float4*>(sample_dev));
    return check_launch("k_sample_velocities");
}

size_t hx_msd_scratch_bytes(unsigned nlags) { return static_cast<size_t>(nlags) * MSD_BLOCKS * sizeof(wf); }

int hx_msd_block(const hx_float4* first_dev, const hx_float4* const* samples_dev, unsigned nlags, size_t nparticle,
                 double* out_dev, void* scratch_dev, size_t scratch_bytes, void* stream)
{
    return hx_correlate_block(HX_CORR_MSD, first_dev, samples_dev, nlags, nparticle, out_dev, scratch_dev, scratch_bytes, stream);
}

int hx_correlate_block(int kind, const hx_float4* first_dev, const hx_float4* const* samples_dev, unsigned nlags,
                       size_t nparticle, double* out_dev, void* scratch_dev, size_t scratch_bytes, void* stream)
{
    if (!first_dev || !samples_dev || !out_dev || !scratch_dev || nlags == 0 || nlags > 65535
        || scratch_bytes < hx_msd_scratch_bytes(nlags) || nparticle >= 0x7FFFFFFFull || kind < HX_CORR_MSD
        || kind > HX_CORR_VACF) {
        set_error("hx_correlate_block: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    wf* partial = static_cast<wf*>(scratch_dev);
    dim3 const grid(MSD_BLOCKS, nlags);
    const float4* f = reinterpret_cast<const float4*>(first_dev);
    const float4* const* sp = reinterpret_cast<const float4* const*>(samples_dev);
    unsigned const n = static_cast<unsigned>(nparticle);
    if (kind == HX_CORR_MSD) k_msd_partial<HX_CORR_MSD><<<grid, MSD_THREADS, 0, st>>>(f, sp, n, partial);
    else if (kind == HX_CORR_MQD) k_msd_partial<HX_CORR_MQD><<<grid, MSD_THREADS, 0, st>>>(f, sp, n, partial);
    else k_msd_partial<HX_CORR_VACF><<<grid, MSD_THREADS, 0, st>>>(f, sp, n, partial);
    int rc = check_launch("k_msd_partial");
    if (rc) return rc;
    k_msd_final<<<nlags, 32, 0, st>>>(partial, MSD_BLOCKS, out_dev);
    return check_launch("k_msd_final");
}

int hx_density_modes(const hx_float4* position_hi_dev, size_t nparticle, const float* q_dev, unsigned nq,
                     double* rho_dev, void* stream)
{
    if (!position_hi_dev || !q_dev || !rho_dev || nq == 0 || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_density_modes: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err = cudaMemsetAsync(rho_dev, 0, 2 * static_cast<size_t>(nq) * sizeof(double), st);
    if (err != cudaSuccess) { set_error("hx_density_modes: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    if (nparticle == 0) return HX_OK;
    unsigned bx = div_up(nparticle, DM_THREADS);
    if (bx > 296) bx = 296;
    k_density_modes<<<dim3(bx, div_up(nq, DM_QTILE)), DM_THREADS, 0, st>>>(
        reinterpret_cast<const float4*>(position_hi_dev), static_cast<unsigned>(nparticle), q_dev, nq, rho_dev);
    int const rc = check_launch("k_density_modes");
    if (rc) return rc;
    k_density_modes_finalise<<<div_up(2 * static_cast<size_t>(nq), 256), 256, 0, st>>>(rho_dev, 2 * nq);
    return check_launch("k_density_modes_finalise");
}

int hx_thermodynamics(const hx_float4* velocity_hi_dev, const float* en_pot_dev, const float* stress_dev,
                      size_t stress_stride, size_t nparticle, double* out_dev, void* stream)
{
    if (!velocity_hi_dev || !out_dev || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_thermodynamics: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    cudaError_t err = cudaMemsetAsync(out_dev, 0, 8 * sizeof(double), st);
    if (err != cudaSuccess) { set_error("hx_thermodynamics: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    if (nparticle == 0) return HX_OK;
    unsigned blocks = div_up(nparticle, TH_THREADS);
    if (blocks > TH_BLOCKS) blocks = TH_BLOCKS;
    k_thermodynamics<<<blocks, TH_THREADS, 0, st>>>(reinterpret_cast<const float4*>(velocity_hi_dev), en_pot_dev,
                                                   stress_dev, stress_stride, static_cast<unsigned>(nparticle), out_dev);
    return check_launch("k_thermodynamics");
}

} // extern "C"
