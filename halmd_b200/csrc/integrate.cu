// Particle state initialisation, velocity-Verlet (double-single), max-displacement reduction,
// rand48 and the Andersen thermostat.
//
// Reference: halmd/mdsim/gpu/particle.cpp:93-140 (initial values),
// halmd/mdsim/gpu/integrators/verlet_kernel.cu:37-98, verlet_nvt_andersen_kernel.cu:45-171,
// halmd/mdsim/gpu/max_displacement_kernel.cu:36-68, halmd/random/gpu/rand48_kernel.{cu,cuh},
// halmd/random/gpu/normal_distribution.cuh:49-67, halmd/numeric/mp/uint48.hpp:72-111.
#include "common.cuh"

namespace hx {

// ---- state init ---------------------------------------------------------------------------------------
__global__ void k_state_init(float4* pos, float4* vel, float4* image, float4* force, unsigned* id, unsigned* rid,
                             unsigned n, unsigned A)
{
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    // position (0,0,0) species 0 — ghosts carry species -1U; velocity 0, mass 1 (particle.cpp:93-113)
    pos[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(i < n ? 0u : HX_PH));
    pos[i + A] = make_float4(0.f, 0.f, 0.f, 0.f);
    vel[i] = make_float4(0.f, 0.f, 0.f, 1.f);
    vel[i + A] = make_float4(0.f, 0.f, 0.f, 0.f);
    image[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    force[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    id[i] = i;
    rid[i] = i;
}

// ---- velocity-Verlet ---------------------------------------------------------------------------------
// (ds3, ds_add_prod, half_kick, drift_wrap: common.cuh — shared with the fused tail of the shell force kernel)

// 16-byte load that the compiler keeps in program order (ahead of the halted test below)
__device__ __forceinline__ float4 ld_ordered(const float4* p)
{
    float4 v;
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// pos_out: where the new positions go — pos itself in the single-GPU loop; the other buffer of the double-buffered
// positions in the decomposed loop, whose ghost slots the neighbour ranks fill over NVLink while this rank's force kernel
// may still read the old buffer (dd.cu).  All loads of pos precede the stores, so the two may alias.
template <bool FINALIZE, bool INTEGRATE>
#ifndef V_THREADS_N
#define V_THREADS_N 128
#endif
__global__ void __launch_bounds__(V_THREADS_N) k_verlet(const float4* pos, float4* pos_out, float4* __restrict__ image,
                                               float4* __restrict__ vel, const float4* __restrict__ force,
                                               unsigned n, unsigned nproc, unsigned A, float dt, vec3 L, vec3 invL,
                                               const float4* __restrict__ r0, unsigned* __restrict__ ctl)
{
    pdl_wait();
    pdl_launch();
    // All loads of the step are issued before the halted flag is tested and before any arithmetic: one memory
    // round trip per thread instead of four dependent ones (the kernel is latency-bound at 256 k particles).
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    bool const act = i < nproc;
    bool const disp = INTEGRATE && r0 && i < n;
    float4 vh, vl, f, ph, pl, q;
    vh = vl = f = ph = pl = q = make_float4(0.f, 0.f, 0.f, 1.f);
    if (act) {
        vh = ld_ordered(vel + i); vl = ld_ordered(vel + i + A);
        f = ld_ordered(force + i);
        if (INTEGRATE) { ph = ld_ordered(pos + i); pl = ld_ordered(pos + i + A); }
        if (disp) q = ld_ordered(r0 + i);
    }
    unsigned halted = 0u;
    if (ctl) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(halted) : "l"(ctl + CTL_HALTED));
    if (halted) return;
    float rr = 0.f;
    if (act) {
        float const mass = vh.w;
        float const half_dt = __fmul_rn(dt, 0.5f);      // timestep / 2 in float (verlet_kernel.cu:59)
        ds3 v = {make_ds(vh.x, vl.x), make_ds(vh.y, vl.y), make_ds(vh.z, vl.z)};
        if (FINALIZE) half_kick(v, f, half_dt, mass);
        if (INTEGRATE) {
            half_kick(v, f, half_dt, mass);
            ds rx = make_ds(ph.x, pl.x), ry = make_ds(ph.y, pl.y), rz = make_ds(ph.z, pl.z);
            float ix = drift_wrap(rx, v.x, dt, L.x);
            float iy = drift_wrap(ry, v.y, dt, L.y);
            float iz = drift_wrap(rz, v.z, dt, L.z);
            pos_out[i] = make_float4(rx.hi, ry.hi, rz.hi, ph.w);
            pos_out[i + A] = make_float4(rx.lo, ry.lo, rz.lo, 0.f);
            if (!(ix == 0.f && iy == 0.f && iz == 0.f)) {
                float4 im = image[i];
                image[i] = make_float4(__fadd_rn(ix, im.x), __fadd_rn(iy, im.y), __fadd_rn(iz, im.z), 0.f);
            }
            if (disp) {
                // max_displacement_kernel.cu:50-58 on the new high words
                float dx = minimg(__fsub_rn(rx.hi, q.x), L.x, invL.x);
                float dy = minimg(__fsub_rn(ry.hi, q.y), L.y, invL.y);
                float dz = minimg(__fsub_rn(rz.hi, q.z), L.z, invL.z);
                rr = norm2(dx, dy, dz);
            }
        }
        vel[i] = make_float4(v.x.hi, v.y.hi, v.z.hi, mass);
        vel[i + A] = make_float4(v.x.lo, v.y.lo, v.z.lo, 0.f);
    }
    if (INTEGRATE && r0) {
        rr = warp_max(rr);
        // One atomic per warp on a single address serialises in L2 (250 k warps at 8 M particles: ~0.5 ms).  The
        // maximum only grows, so a warp whose value does not exceed the maximum it can read (a possibly stale, hence
        // smaller, copy) has nothing to contribute; after the first warps almost all skip the atomic.
        if ((threadIdx.x & 31u) == 0 && rr > 0.f && __float_as_uint(rr) > __ldcg(&ctl[CTL_MAX_RR]))
            atomicMax(&ctl[CTL_MAX_RR], __float_as_uint(rr));
    }
}

// ---- max displacement (stand-alone) --------------------------------------------------------------------
__global__ void k_displacement(const float4* __restrict__ pos, const float4* __restrict__ r0, unsigned n, vec3 L,
                               vec3 invL, unsigned* __restrict__ max_rr_bits)
{
    float rr = 0.f;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pos[i], q = r0[i];
        float dx = minimg(__fsub_rn(p.x, q.x), L.x, invL.x);
        float dy = minimg(__fsub_rn(p.y, q.y), L.y, invL.y);
        float dz = minimg(__fsub_rn(p.z, q.z), L.z, invL.z);
        rr = fmaxf(rr, norm2(dx, dy, dz));
    }
    rr = warp_max(rr);
    if ((threadIdx.x & 31u) == 0 && rr > 0.f && __float_as_uint(rr) > __ldcg(max_rr_bits))
        atomicMax(max_rr_bits, __float_as_uint(rr));
}

// ---- rand48 ---------------------------------------------------------------------------------------------
constexpr unsigned long long R48_MASK = (1ull << 48) - 1;
constexpr unsigned long long R48_A = 0x5DEECE66Dull;
constexpr unsigned long long R48_C = 0xBull;

__host__ __device__ inline unsigned long long r48_pack(const uint16_t* s)
{
    return static_cast<unsigned long long>(s[0]) | (static_cast<unsigned long long>(s[1]) << 16)
           | (static_cast<unsigned long long>(s[2]) << 32);
}

__host__ __device__ inline void r48_unpack(unsigned long long v, uint16_t* s)
{
    s[0] = static_cast<uint16_t>(v & 0xFFFF);
    s[1] = static_cast<uint16_t>((v >> 16) & 0xFFFF);
    s[2] = static_cast<uint16_t>((v >> 32) & 0xFFFF);
}

// thread t starts at x_{t+1} = a^(t+1) x_0 + c (1 + a + ... + a^t)   (rand48_kernel.cu:42-100)
__global__ void k_rand48_seed(uint16_t* __restrict__ state, unsigned nthreads, unsigned long long x0)
{
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    // (A, C) of the affine map applied t+1 times, by binary exponentiation of (a, c)
    unsigned long long A = 1, C = 0;           // identity
    unsigned long long a = R48_A, c = R48_C;   // current power
    for (unsigned k = t + 1; k > 0; k >>= 1) {
        if (k & 1u) {
            // compose: x -> a (A x + C) + c
            A = (a * A) & R48_MASK;
            C = (a * C + c) & R48_MASK;
        }
        // square: x -> a (a x + c) + c
        c = (a * c + c) & R48_MASK;
        a = (a * a) & R48_MASK;
    }
    r48_unpack((A * x0 + C) & R48_MASK, state + 3 * t);
}

struct r48 {
    unsigned long long s, a, c;
    // rand48_kernel.cuh:59-65 — variate from the state before advancing
    __device__ __forceinline__ float uniform()
    {
        float z = static_cast<float>((s >> 32) & 0xFFFF), y = static_cast<float>((s >> 16) & 0xFFFF);
        float u = __fadd_rn(__fmul_rn(z, 1.f / 65536.f), __fmul_rn(y, 1.f / 4294967296.f));
        s = (a * s + c) & R48_MASK;
        return u;
    }
};

__global__ void k_rand48_uniform(uint16_t* __restrict__ state, unsigned long long a, unsigned long long c,
                                 float* __restrict__ out, unsigned n)
{
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    r48 g = {r48_pack(state + 3 * t), a, c};
    for (unsigned i = t; i < n; i += gridDim.x * blockDim.x) out[i] = g.uniform();
    r48_unpack(g.s, state + 3 * t);
}

// Box–Muller polar method (normal_distribution.cuh:49-67); the rejection test is evaluated in a fixed
// operation order so that the CPU oracle consumes the same number of variates.
__device__ __forceinline__ void normal_pair(r48& g, float sigma, float& n1, float& n2)
{
    float v1, v2, s;
    do {
        v1 = __fmaf_rn(2.f, g.uniform(), -1.f);
        v2 = __fmaf_rn(2.f, g.uniform(), -1.f);
        s = __fmaf_rn(v1, v1, __fmul_rn(v2, v2));
    } while (s >= 1.f);
    float w = sigma * sqrtf(-2.f * logf(s) / s);
    n1 = w * v1;
    n2 = w * v2;
}

// verlet_nvt_andersen_kernel.cu:96-171 — launched with the RNG's grid (blocks x threads); the warp-wide
// collision decision and the grid-stride particle mapping are part of the algorithm's definition.
__global__ void k_andersen_finalize(float4* __restrict__ vel, const float4* __restrict__ force, unsigned n, unsigned A,
                                    float dt, float sqrt_temperature, float q, uint16_t* __restrict__ state,
                                    unsigned long long ra, unsigned long long rc, const unsigned* __restrict__ ctl)
{
    if (ctl && ctl[CTL_HALTED]) return;
    unsigned const t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned const T = gridDim.x * blockDim.x;
    r48 g = {r48_pack(state + 3 * t), ra, rc};
    bool cached = false;
    float cache = 0.f;
    float const half_dt = __fmul_rn(dt, 0.5f);
    // all lanes of a warp must take part in the vote: iterate over whole warps
    unsigned const warp_base = t & ~31u;
    for (unsigned base = warp_base; base < n; base += T) {
        unsigned i = base + (t & 31u);
        bool const valid = i < n;
        // lanes beyond n left the loop in the reference (`i < npart`); they do not vote and draw nothing
        unsigned const active = __ballot_sync(0xFFFFFFFFu, valid);
        if (!valid) continue;
        float4 vh = vel[i], vl = vel[i + A];
        float const mass = vh.w;
        bool const no_coll = g.uniform() > q;
        if (__all_sync(active, no_coll)) {
            float4 f = force[i];
            ds3 v = {make_ds(vh.x, vl.x), make_ds(vh.y, vl.y), make_ds(vh.z, vl.z)};
            half_kick(v, f, half_dt, mass);
            vel[i] = make_float4(v.x.hi, v.y.hi, v.z.hi, mass);
            vel[i + A] = make_float4(v.x.lo, v.y.lo, v.z.lo, 0.f);
        } else {
            float vx, vy, vz;
            normal_pair(g, sqrt_temperature, vx, vy);
            if ((cached = !cached)) {
                normal_pair(g, sqrt_temperature, vz, cache);
            } else {
                vz = cache;
            }
            vel[i] = make_float4(vx, vy, vz, mass);
            vel[i + A] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    r48_unpack(g.s, state + 3 * t);
}

// Double-single product with both low words live, contracted as nvcc --fmad=true contracts dsmul (dsfun.hpp:178-207)
// in the reference build: the Dekker splits and the partial products ride on FFMA, only a0 * b0 is rounded on its own.
__device__ __forceinline__ ds ds_mul_contracted(ds a, ds b)
{
    float const sa1 = __fmaf_rn(a.hi, 8193.0f, -__fmaf_rn(a.hi, 8193.0f, -a.hi));
    float const sb1 = __fmaf_rn(b.hi, 8193.0f, -__fmaf_rn(b.hi, 8193.0f, -b.hi));
    float const sa2 = __fsub_rn(a.hi, sa1);
    float const sb2 = __fsub_rn(b.hi, sb1);
    float const c11 = __fmul_rn(a.hi, b.hi);
    float c21 = __fmaf_rn(sa1, sb1, -c11);
    c21 = __fmaf_rn(sa1, sb2, c21);
    c21 = __fmaf_rn(sb1, sa2, c21);
    c21 = __fmaf_rn(sa2, sb2, c21);
    float const c2 = __fmaf_rn(a.hi, b.lo, __fmul_rn(a.lo, b.hi));
    float const t1 = __fadd_rn(c11, c2);
    float const e = __fsub_rn(t1, c11);
    float t2 = __fadd_rn(__fsub_rn(c2, e), __fsub_rn(c11, __fsub_rn(t1, e)));
    t2 = __fmaf_rn(a.lo, b.lo, __fadd_rn(c21, t2));
    ds c;
    c.hi = __fadd_rn(t1, t2);
    c.lo = __fsub_rn(t2, __fsub_rn(c.hi, t1));
    return c;
}

// ---- Nose-Hoover chain (verlet_nvt_hoover_kernel.cu:37-125) -------------------------------------------------------------
// integrate: v *= scale (double-single product), then the velocity-Verlet first half-step; rescale: v *= scale.
// The chain variables live on the host (verlet_nvt_hoover.cpp:185-222); finalize is the plain Verlet half-kick.
__global__ void __launch_bounds__(128) k_hoover_integrate(float4* __restrict__ pos, float4* __restrict__ image,
                                                         float4* __restrict__ vel, const float4* __restrict__ force,
                                                         unsigned A, float dt, ds scale, vec3 L)
{
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    float4 vh = vel[i], vl = vel[i + A];
    float4 const f = force[i];
    float const mass = vh.w;
    ds3 v = {ds_mul_contracted(make_ds(vh.x, vl.x), scale), ds_mul_contracted(make_ds(vh.y, vl.y), scale),
             ds_mul_contracted(make_ds(vh.z, vl.z), scale)};
    half_kick(v, f, __fmul_rn(dt, 0.5f), mass);
    float4 const ph = pos[i], pl = pos[i + A];
    ds rx = make_ds(ph.x, pl.x), ry = make_ds(ph.y, pl.y), rz = make_ds(ph.z, pl.z);
    float const ix = drift_wrap(rx, v.x, dt, L.x);
    float const iy = drift_wrap(ry, v.y, dt, L.y);
    float const iz = drift_wrap(rz, v.z, dt, L.z);
    pos[i] = make_float4(rx.hi, ry.hi, rz.hi, ph.w);
    pos[i + A] = make_float4(rx.lo, ry.lo, rz.lo, 0.f);
    vel[i] = make_float4(v.x.hi, v.y.hi, v.z.hi, mass);
    vel[i + A] = make_float4(v.x.lo, v.y.lo, v.z.lo, 0.f);
    if (!(ix == 0.f && iy == 0.f && iz == 0.f)) {
        float4 const im = image[i];
        image[i] = make_float4(__fadd_rn(ix, im.x), __fadd_rn(iy, im.y), __fadd_rn(iz, im.z), 0.f);
    }
}

__global__ void __launch_bounds__(128) k_hoover_rescale(float4* __restrict__ vel, unsigned A, ds scale)
{
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A) return;
    float4 const vh = vel[i], vl = vel[i + A];
    ds const x = ds_mul_contracted(make_ds(vh.x, vl.x), scale), y = ds_mul_contracted(make_ds(vh.y, vl.y), scale),
             z = ds_mul_contracted(make_ds(vh.z, vl.z), scale);
    vel[i] = make_float4(x.hi, y.hi, z.hi, vh.w);
    vel[i + A] = make_float4(x.lo, y.lo, z.lo, 0.f);
}

// ---- Andersen thermostat in two passes --------------------------------------------------------------------------------
// The reference kernel above is sequential by construction: 8192 generator threads walk the particle array with a
// grid stride, and what a thread draws at one particle depends on the collisions of its warp at all earlier ones
// (warp-wide vote, cached second variate, rejection loops of the polar method).  At 10^6 particles that is 122
// dependent iterations of memory traffic, double-single arithmetic, log and sqrt on a quarter of a percent of the
// machine (100 us where the memory traffic needs 12).  The same stream is reproduced in two passes:
//   plan     (generator grid, sequential per thread): only the generator bookkeeping — the uniform variate of the
//            collision test, the warp vote, the rejection loops — recording, per particle-warp, whether it collides
//            and, per colliding particle, the generator states at which its accepted variate pairs start;
//   apply    (one thread per particle): half-kick, or Box-Muller from the recorded states.
// Same functions on the same states: velocities and generator state are bit-identical to the one-pass kernel
// (tests/test_gpu_kernels.py::test_andersen_two_pass_equals_one_pass).
__device__ __forceinline__ unsigned long long accepted_pair_state(r48& g)
{
    unsigned long long s0;
    float s;
    do {
        s0 = g.s;
        float v1 = __fmaf_rn(2.f, g.uniform(), -1.f);
        float v2 = __fmaf_rn(2.f, g.uniform(), -1.f);
        s = __fmaf_rn(v1, v1, __fmul_rn(v2, v2));
    } while (s >= 1.f);
    return s0;
}

__global__ void k_andersen_plan(unsigned n, float q, uint16_t* __restrict__ state, unsigned long long ra, unsigned long long rc,
                                unsigned char* __restrict__ collide, unsigned long long* __restrict__ plan_xy,
                                unsigned long long* __restrict__ plan_z, const unsigned* __restrict__ ctl)
{
    pdl_wait();
    pdl_launch();
    if (ctl && ctl[CTL_HALTED]) return;
    unsigned const t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned const T = gridDim.x * blockDim.x;
    r48 g = {r48_pack(state + 3 * t), ra, rc};
    bool cached = false;
    unsigned long long last_z = 0;
    unsigned const warp_base = t & ~31u;
    for (unsigned base = warp_base; base < n; base += T) {
        unsigned i = base + (t & 31u);
        bool const valid = i < n;
        unsigned const active = __ballot_sync(0xFFFFFFFFu, valid);
        if (!valid) continue;
        bool const no_coll = g.uniform() > q;
        bool const det = __all_sync(active, no_coll);
        if ((t & 31u) == 0) collide[base >> 5] = det ? 0 : 1;
        if (!det) {
            plan_xy[i] = accepted_pair_state(g);
            if ((cached = !cached)) {
                last_z = accepted_pair_state(g);
                plan_z[i] = last_z;                         // first variate of a new pair
            } else {
                plan_z[i] = last_z | (1ull << 63);          // cached second variate of the previous pair
            }
        }
    }
    r48_unpack(g.s, state + 3 * t);
}

__global__ void __launch_bounds__(128) k_andersen_apply(float4* __restrict__ vel, const float4* __restrict__ force, unsigned n,
                                                       unsigned A, float dt, float sqrt_temperature, unsigned long long ra,
                                                       unsigned long long rc, const unsigned char* __restrict__ collide,
                                                       const unsigned long long* __restrict__ plan_xy,
                                                       const unsigned long long* __restrict__ plan_z,
                                                       const unsigned* __restrict__ ctl)
{
    pdl_wait();
    pdl_launch();
    if (ctl && ctl[CTL_HALTED]) return;
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 const vh = vel[i];
    float const mass = vh.w;
    if (!collide[i >> 5]) {
        float4 const vl = vel[i + A];
        float4 const f = force[i];
        ds3 v = {make_ds(vh.x, vl.x), make_ds(vh.y, vl.y), make_ds(vh.z, vl.z)};
        half_kick(v, f, __fmul_rn(dt, 0.5f), mass);
        vel[i] = make_float4(v.x.hi, v.y.hi, v.z.hi, mass);
        vel[i + A] = make_float4(v.x.lo, v.y.lo, v.z.lo, 0.f);
    } else {
        float vx, vy, z1, z2;
        r48 gx = {plan_xy[i], ra, rc};
        normal_pair(gx, sqrt_temperature, vx, vy);
        unsigned long long const pz = plan_z[i];
        r48 gz = {pz & R48_MASK, ra, rc};
        normal_pair(gz, sqrt_temperature, z1, z2);
        vel[i] = make_float4(vx, vy, (pz >> 63) ? z2 : z1, mass);
        vel[i + A] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

} // namespace hx

using namespace hx;

extern "C" {

int hx_state_init(hx_float4* position_dev, hx_float4* velocity_dev, hx_float4* image_dev, hx_float4* force_dev,
                  uint32_t* id_dev, uint32_t* reverse_id_dev, size_t nparticle, size_t array_size, void* stream)
{
    if (!position_dev || !velocity_dev || !image_dev || !force_dev || !id_dev || !reverse_id_dev
        || nparticle > array_size || array_size == 0 || array_size >= 0x7FFFFFFFull) {
        set_error("hx_state_init: invalid argument");
        return HX_E_INVALID;
    }
    k_state_init<<<div_up(array_size, 256), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<float4*>(position_dev), reinterpret_cast<float4*>(velocity_dev),
        reinterpret_cast<float4*>(image_dev), reinterpret_cast<float4*>(force_dev), id_dev, reverse_id_dev,
        static_cast<unsigned>(nparticle), static_cast<unsigned>(array_size));
    return check_launch("k_state_init");
}

} // extern "C"

namespace hx {

int verlet_launch(bool fin, bool integ, hx_float4* pos, hx_float4* image, hx_float4* vel,
                  const hx_float4* force, size_t n, size_t A, float dt, const float L[3],
                  const hx_float4* r0, uint32_t* ctl, void* stream, size_t nproc_arg, hx_float4* pos_out_arg)
{
    size_t const nproc = nproc_arg == ~size_t(0) ? A : nproc_arg;
    if (!vel || !force || A == 0 || A >= 0x7FFFFFFFull || n > A || nproc > A || (integ && (!pos || !image || !L))
        || (r0 && !ctl)) {
        set_error("hx_verlet: invalid argument");
        return HX_E_INVALID;
    }
    vec3 Lv = {1.f, 1.f, 1.f}, iLv = {1.f, 1.f, 1.f};
    if (L) {
        Lv = make_vec3(L);
        iLv = {1.0f / L[0], 1.0f / L[1], 1.0f / L[2]};
    }
    dim3 grid(div_up(nproc > 0 ? nproc : 1, V_THREADS_N)), block(V_THREADS_N);
    cudaStream_t st = as_stream(stream);
    float4* p = reinterpret_cast<float4*>(pos);
    float4* po = reinterpret_cast<float4*>(pos_out_arg ? pos_out_arg : pos);
    float4* im = reinterpret_cast<float4*>(image);
    float4* v = reinterpret_cast<float4*>(vel);
    const float4* f = reinterpret_cast<const float4*>(force);
    const float4* q = reinterpret_cast<const float4*>(r0);
    unsigned nn = static_cast<unsigned>(n), AA = static_cast<unsigned>(A), np = static_cast<unsigned>(nproc);
    if (fin && integ) launch_pdl(k_verlet<true, true>, grid, block, 0, st, p, po, im, v, f, nn, np, AA, dt, Lv, iLv, q, ctl);
    else if (integ) launch_pdl(k_verlet<false, true>, grid, block, 0, st, p, po, im, v, f, nn, np, AA, dt, Lv, iLv, q, ctl);
    else launch_pdl(k_verlet<true, false>, grid, block, 0, st, p, po, im, v, f, nn, np, AA, dt, Lv, iLv, q, ctl);
    return check_launch("k_verlet");
}

} // namespace hx

extern "C" {

int hx_verlet_integrate(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev,
                        const hx_float4* force_dev, size_t array_size, float timestep, const float box_length[3],
                        void* stream)
{
    return verlet_launch(false, true, position_dev, image_dev, velocity_dev, force_dev, array_size, array_size,
                         timestep, box_length, nullptr, nullptr, stream);
}

int hx_verlet_finalize(hx_float4* velocity_dev, const hx_float4* force_dev, size_t array_size, float timestep,
                       void* stream)
{
    return verlet_launch(true, false, nullptr, nullptr, velocity_dev, force_dev, array_size, array_size, timestep,
                         nullptr, nullptr, nullptr, stream);
}

int hx_verlet_finalize_integrate(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev,
                                 const hx_float4* force_dev, size_t nparticle, size_t array_size, float timestep,
                                 const float box_length[3], int do_finalize, const hx_float4* r0_dev,
                                 uint32_t* ctl_dev, void* stream)
{
    return verlet_launch(do_finalize != 0, true, position_dev, image_dev, velocity_dev, force_dev, nparticle,
                         array_size, timestep, box_length, r0_dev, ctl_dev, stream);
}

int hx_verlet_finalize_integrate_range(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev,
                                       const hx_float4* force_dev, size_t nparticle, size_t nprocess, size_t array_size,
                                       float timestep, const float box_length[3], int do_finalize, int do_integrate,
                                       const hx_float4* r0_dev, uint32_t* ctl_dev, void* stream)
{
    if (!do_finalize && !do_integrate) return HX_OK;
    return verlet_launch(do_finalize != 0, do_integrate != 0, position_dev, image_dev, velocity_dev, force_dev, nparticle,
                         array_size, timestep, box_length, do_integrate ? r0_dev : nullptr, ctl_dev, stream, nprocess);
}

int hx_verlet_finalize_guarded(hx_float4* velocity_dev, const hx_float4* force_dev, size_t array_size,
                               float timestep, uint32_t* ctl_dev, void* stream)
{
    return verlet_launch(true, false, nullptr, nullptr, velocity_dev, force_dev, array_size, array_size, timestep,
                         nullptr, nullptr, ctl_dev, stream);
}

int hx_ctl_reset(uint32_t* ctl_dev, void* stream)
{
    if (!ctl_dev) { set_error("hx_ctl_reset: null pointer"); return HX_E_INVALID; }
    cudaError_t err = cudaMemsetAsync(ctl_dev, 0, 4 * sizeof(uint32_t), as_stream(stream));
    if (err != cudaSuccess) { set_error("hx_ctl_reset: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    return HX_OK;
}

int hx_displacement_zero(const hx_float4* position_hi_dev, hx_float4* r0_dev, size_t nparticle, void* stream)
{
    if (!position_hi_dev || !r0_dev) { set_error("hx_displacement_zero: null pointer"); return HX_E_INVALID; }
    cudaError_t err = cudaMemcpyAsync(r0_dev, position_hi_dev, nparticle * sizeof(hx_float4),
                                      cudaMemcpyDeviceToDevice, as_stream(stream));
    if (err != cudaSuccess) { set_error("hx_displacement_zero: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    return HX_OK;
}

int hx_displacement_max(const hx_float4* position_hi_dev, const hx_float4* r0_dev, size_t nparticle,
                        const float box_length[3], uint32_t* max_rr_bits_dev, int reset, void* stream)
{
    if (!position_hi_dev || !r0_dev || !box_length || !max_rr_bits_dev || nparticle >= 0x7FFFFFFFull) {
        set_error("hx_displacement_max: invalid argument");
        return HX_E_INVALID;
    }
    cudaStream_t st = as_stream(stream);
    if (reset) {
        cudaError_t err = cudaMemsetAsync(max_rr_bits_dev, 0, sizeof(uint32_t), st);
        if (err != cudaSuccess) { set_error("hx_displacement_max: %s", cudaGetErrorString(err)); return static_cast<int>(err); }
    }
    if (nparticle == 0) return HX_OK;
    vec3 L = make_vec3(box_length);
    vec3 iL = {1.0f / box_length[0], 1.0f / box_length[1], 1.0f / box_length[2]};
    unsigned blocks = div_up(nparticle, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_displacement<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(position_hi_dev),
                                          reinterpret_cast<const float4*>(r0_dev), static_cast<unsigned>(nparticle),
                                          L, iL, max_rr_bits_dev);
    return check_launch("k_displacement");
}

int hx_rand48_seed(hx_rand48* rng, unsigned blocks, unsigned threads, unsigned seed, uint16_t* state_dev,
                   void* stream)
{
    if (!rng || !state_dev || blocks == 0 || threads == 0 || threads % 32 != 0 || threads > 1024) {
        set_error("hx_rand48_seed: invalid argument");
        return HX_E_INVALID;
    }
    unsigned const T = blocks * threads;
    rng->blocks = blocks;
    rng->threads = threads;
    // leapfrog constants: the affine map applied T times (rand48_kernel.cu:78-82, thread GTDIM-1)
    unsigned long long A = 1, C = 0, a = R48_A, c = R48_C;
    for (unsigned k = T; k > 0; k >>= 1) {
        if (k & 1u) {
            A = (a * A) & R48_MASK;
            C = (a * C + c) & R48_MASK;
        }
        c = (a * c + c) & R48_MASK;
        a = (a * a) & R48_MASK;
    }
    r48_unpack(A, rng->a);
    r48_unpack(C, rng->c);
    // default seed words (0x330E, 0xABCD, 0x1234); a non-zero seed replaces the upper two (rand48_kernel.cu:85-92)
    uint16_t x[3] = {0x330E, 0xABCD, 0x1234};
    if (seed > 0) {
        x[1] = seed & 0xFFFF;
        x[2] = (seed >> 16) & 0xFFFF;
    }
    k_rand48_seed<<<blocks, threads, 0, as_stream(stream)>>>(state_dev, T, r48_pack(x));
    return check_launch("k_rand48_seed");
}

int hx_rand48_uniform(const hx_rand48* rng, uint16_t* state_dev, float* out_dev, size_t n, void* stream)
{
    if (!rng || !state_dev || !out_dev || n >= 0x7FFFFFFFull) {
        set_error("hx_rand48_uniform: invalid argument");
        return HX_E_INVALID;
    }
    k_rand48_uniform<<<rng->blocks, rng->threads, 0, as_stream(stream)>>>(state_dev, r48_pack(rng->a), r48_pack(rng->c),
                                                                         out_dev, static_cast<unsigned>(n));
    return check_launch("k_rand48_uniform");
}

static int andersen_launch(hx_float4* velocity_dev, const hx_float4* force_dev, size_t nparticle, size_t array_size,
                           float timestep, float sqrt_temperature, float coll_prob, const hx_rand48* rng,
                           uint16_t* state_dev, const uint32_t* ctl, void* stream)
{
    if (!velocity_dev || !force_dev || !rng || !state_dev || nparticle > array_size || array_size >= 0x7FFFFFFFull) {
        set_error("hx_andersen_finalize: invalid argument");
        return HX_E_INVALID;
    }
    // pow(1 - q, warpSize) = 1 - coll_prob  (verlet_nvt_andersen_kernel.cu:121-123); evaluated on the host in
    // IEEE float instead of the device's __powf approximation
    float q = 1.f - powf(1.f - coll_prob, 1.f / 32);
    k_andersen_finalize<<<rng->blocks, rng->threads, 0, as_stream(stream)>>>(
        reinterpret_cast<float4*>(velocity_dev), reinterpret_cast<const float4*>(force_dev),
        static_cast<unsigned>(nparticle), static_cast<unsigned>(array_size), timestep, sqrt_temperature, q, state_dev,
        r48_pack(rng->a), r48_pack(rng->c), ctl);
    return check_launch("k_andersen_finalize");
}

// double -> double-single as halmd::dsfloat(double) does (numeric/mp/dsfloat.hpp: hi = float(a), lo = float(a - hi))
static ds ds_from_double(double a)
{
    ds r;
    r.hi = static_cast<float>(a);
    r.lo = static_cast<float>(a - static_cast<double>(r.hi));
    return r;
}

int hx_hoover_integrate(hx_float4* position_dev, hx_float4* image_dev, hx_float4* velocity_dev, const hx_float4* force_dev,
                        size_t array_size, float timestep, double scale, const float box_length[3], void* stream)
{
    if (!position_dev || !image_dev || !velocity_dev || !force_dev || !box_length || array_size == 0
        || array_size >= 0x7FFFFFFFull) {
        set_error("hx_hoover_integrate: invalid argument");
        return HX_E_INVALID;
    }
    k_hoover_integrate<<<div_up(array_size, 128), 128, 0, as_stream(stream)>>>(
        reinterpret_cast<float4*>(position_dev), reinterpret_cast<float4*>(image_dev), reinterpret_cast<float4*>(velocity_dev),
        reinterpret_cast<const float4*>(force_dev), static_cast<unsigned>(array_size), timestep, ds_from_double(scale),
        make_vec3(box_length));
    return check_launch("k_hoover_integrate");
}

int hx_hoover_rescale(hx_float4* velocity_dev, size_t array_size, double scale, void* stream)
{
    if (!velocity_dev || array_size == 0 || array_size >= 0x7FFFFFFFull) {
        set_error("hx_hoover_rescale: invalid argument");
        return HX_E_INVALID;
    }
    k_hoover_rescale<<<div_up(array_size, 128), 128, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(velocity_dev),
                                                                            static_cast<unsigned>(array_size),
                                                                            ds_from_double(scale));
    return check_launch("k_hoover_rescale");
}

size_t hx_andersen_scratch_bytes(size_t nparticle)
{
    // collision flag per particle-warp (padded to 8 bytes) + two generator states per particle
    return ((nparticle + 31) / 32 + 7) / 8 * 8 + 2 * sizeof(unsigned long long) * nparticle;
}

int hx_andersen_finalize2(hx_float4* velocity_dev, const hx_float4* force_dev, size_t nparticle, size_t array_size,
                          float timestep, float sqrt_temperature, float coll_prob, const hx_rand48* rng,
                          uint16_t* state_dev, void* scratch_dev, size_t scratch_bytes, const uint32_t* ctl_dev, void* stream)
{
    if (!velocity_dev || !force_dev || !rng || !state_dev || !scratch_dev || nparticle > array_size
        || array_size >= 0x7FFFFFFFull || scratch_bytes < hx_andersen_scratch_bytes(nparticle)) {
        set_error("hx_andersen_finalize2: invalid argument");
        return HX_E_INVALID;
    }
    if (nparticle == 0) return HX_OK;
    float const q = 1.f - powf(1.f - coll_prob, 1.f / 32);      // as andersen_launch
    size_t const flag_bytes = ((nparticle + 31) / 32 + 7) / 8 * 8;
    unsigned char* collide = static_cast<unsigned char*>(scratch_dev);
    unsigned long long* plan_xy = reinterpret_cast<unsigned long long*>(collide + flag_bytes);
    unsigned long long* plan_z = plan_xy + nparticle;
    unsigned const n = static_cast<unsigned>(nparticle);
    cudaStream_t st = as_stream(stream);
    // (programmatic dependent launches like the rest of the step loop: launch latency overlaps the predecessor's tail)
    launch_pdl(k_andersen_plan, dim3(rng->blocks), dim3(rng->threads), 0, st, n, q, state_dev, r48_pack(rng->a), r48_pack(rng->c),
               collide, plan_xy, plan_z, ctl_dev);
    int rc = check_launch("k_andersen_plan");
    if (rc != HX_OK) return rc;
    launch_pdl(k_andersen_apply, dim3(div_up(nparticle, 128)), dim3(128), 0, st, reinterpret_cast<float4*>(velocity_dev),
               reinterpret_cast<const float4*>(force_dev), n, static_cast<unsigned>(array_size), timestep, sqrt_temperature,
               r48_pack(rng->a), r48_pack(rng->c), collide, plan_xy, plan_z, ctl_dev);
    return check_launch("k_andersen_apply");
}

int hx_andersen_finalize(hx_float4* velocity_dev, const hx_float4* force_dev, size_t nparticle, size_t array_size,
                         float timestep, float sqrt_temperature, float coll_prob, const hx_rand48* rng,
                         uint16_t* state_dev, void* stream)
{
    return andersen_launch(velocity_dev, force_dev, nparticle, array_size, timestep, sqrt_temperature, coll_prob, rng,
                           state_dev, nullptr, stream);
}

int hx_andersen_finalize_guarded(hx_float4* velocity_dev, const hx_float4* force_dev, size_t nparticle,
                                 size_t array_size, float timestep, float sqrt_temperature, float coll_prob,
                                 const hx_rand48* rng, uint16_t* state_dev, const uint32_t* ctl_dev, void* stream)
{
    return andersen_launch(velocity_dev, force_dev, nparticle, array_size, timestep, sqrt_temperature, coll_prob, rng,
                           state_dev, ctl_dev, stream);
}

} // extern "C"
