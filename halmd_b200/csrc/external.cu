// External one-body force: planar walls with a Lennard-Jones 9-3 potential and C2-smooth truncation.
//
// Reference: halmd/mdsim/gpu/forces/external_kernel.cuh:49-103 (one thread per particle: reduce the position into the
// primary cell, evaluate the potential, add to or overwrite the force), potentials/external/planar_wall_kernel.cu:43-105.
// The arithmetic follows what nvcc makes of the reference under -prec-div=false --fmad=true (a / b = a * MUFU.RCP(b)),
// like the pair force (force.cu); the kernel is a pure stream: 16 B read + 16 B (32 B when adding) per particle.
#include "common.cuh"

namespace hx {

struct wall_args {
    hx_planar_wall w;
    const float4* pos;
    float4* force;
    float* en_pot;
    float* stress;
    const unsigned* ctl;
    size_t stress_stride;
    unsigned n;
    vec3 L;
    int force_zero;
};

// a / b as the reference build divides: a * MUFU.RCP(b)
__device__ __forceinline__ float div_approx(float a, float b) { return __fmul_rn(a, rcp_approx(b)); }

template <bool AUX>
__global__ void __launch_bounds__(128) k_planar_wall(wall_args a)
{
    if (a.ctl && a.ctl[CTL_HALTED]) return;
    unsigned const i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    float4 const p = a.pos[i];
    unsigned const species = __float_as_uint(p.w);
    // box_kernel::reduce_periodic (box_kernel.cuh:37-45): r -= rint(r / L) * L with the approximate quotient
    float rx = p.x, ry = p.y, rz = p.z;
    rx = __fmaf_rn(-rintf(div_approx(rx, a.L.x)), a.L.x, rx);
    ry = __fmaf_rn(-rintf(div_approx(ry, a.L.y)), a.L.y, ry);
    rz = __fmaf_rn(-rintf(div_approx(rz, a.L.z)), a.L.z, rz);

    float en_pot = 0.f, fx = 0.f, fy = 0.f, fz = 0.f;
    float const smoothing = a.w.smoothing;
    for (unsigned k = 0; k < a.w.nwall; ++k) {
        float const nx = a.w.surface_normal[k][0], ny = a.w.surface_normal[k][1], nz = a.w.surface_normal[k][2];
        // inner_prod(r, n) - offset; operation order and fused products as in the SASS of the reference kernel
        // (oracle/_ref/obj/planar_wall_kernel.o, external_kernel::compute<false, 3>)
        float const d = __fsub_rn(__fmaf_rn(rz, nz, __fmaf_rn(rx, nx, __fmul_rn(ry, ny))), a.w.offset[k]);
        unsigned const t = species * a.w.nwall + k;
        float const epsilon = a.w.epsilon[t], sigma = a.w.sigma[t], w = a.w.wetting[t], cutoff = a.w.cutoff[t];
        if (fabsf(d) >= cutoff) continue;
        float const k215 = 2.f / 15;
        // cutoff energy: eps dc3i (2/15 dc3i^2 - w)
        float const sc = __fmul_rn(rcp_approx(cutoff), sigma);
        float const dc3i = __fmul_rn(sc, __fmul_rn(sc, sc));
        float const en_cut = __fmul_rn(__fmul_rn(epsilon, dc3i), __fmaf_rn(dc3i, __fmul_rn(dc3i, k215), -w));
        // energy and force
        float const rd = rcp_approx(d);
        float const sd = __fmul_rn(rd, sigma);
        float const d3i = __fmul_rn(sd, __fmul_rn(sd, sd));
        float const d6i = __fmul_rn(d3i, __fmul_rn(d3i, k215));
        float const eps_d3i = __fmul_rn(d3i, d > 0 ? epsilon : -epsilon);
        float en_sub = __fmaf_rn(eps_d3i, __fsub_rn(d6i, w), -en_cut);
        float fval = __fmul_rn(rd, __fmul_rn(__fmul_rn(eps_d3i, 3.f), __fmaf_rn(d6i, 3.f, -w)));
        // smooth truncation
        float const rs = rcp_approx(smoothing);
        float const dd = __fmul_rn(rs, __fsub_rn(fabsf(d), cutoff));
        float const x2 = __fmul_rn(dd, dd);
        float const x4 = __fmul_rn(x2, x2);
        float const x4i = rcp_approx(__fadd_rn(x4, 1.f));
        float const h0 = __fmul_rn(x4, x4i);
        float const nh1 = __fmul_rn(x4i, __fmul_rn(x4i, __fmul_rn(x2, __fmul_rn(dd, -4.f))));      // -h1
        fval = __fmaf_rn(fval, h0, __fmul_rn(rs, __fmul_rn(en_sub, nh1)));
        en_sub = __fmul_rn(h0, en_sub);
        fx = __fmaf_rn(nx, fval, fx);
        fy = __fmaf_rn(ny, fval, fy);
        fz = __fmaf_rn(nz, fval, fz);
        en_pot = __fadd_rn(en_pot, en_sub);
    }
    if (!a.force_zero) {
        float4 const old = a.force[i];
        fx = __fadd_rn(fx, old.x); fy = __fadd_rn(fy, old.y); fz = __fadd_rn(fz, old.z);
    }
    a.force[i] = make_float4(fx, fy, fz, 0.f);
    if (AUX) {
        if (!a.force_zero) {
            en_pot = __fadd_rn(en_pot, a.en_pot[i]);
        } else {
#pragma unroll
            for (int c = 0; c < 6; ++c) a.stress[c * a.stress_stride + i] = 0.f;
        }
        a.en_pot[i] = en_pot;
    }
}

} // namespace hx

using namespace hx;

extern "C" {

int hx_planar_wall_init(hx_planar_wall* w, unsigned nwall, unsigned nspecies, const float* offset, const float* surface_normal,
                        const float* epsilon, const float* sigma, const float* wetting, const float* cutoff, float smoothing)
{
    if (!w || !offset || !surface_normal || !epsilon || !sigma || !wetting || !cutoff || nwall == 0 || nwall > HX_MAX_WALLS
        || nspecies == 0 || nspecies > HX_MAX_WALL_SPECIES || !(smoothing > 0)) {
        set_error("hx_planar_wall_init: invalid argument");
        return HX_E_INVALID;
    }
    *w = hx_planar_wall();
    w->nwall = nwall;
    w->nspecies = nspecies;
    w->smoothing = smoothing;
    for (unsigned i = 0; i < nwall; ++i) {
        // planar_wall.cpp:83-86 — impose normalisation of surface normals (float arithmetic, norm_2 = sqrt(inner_prod))
        const float* n = surface_normal + 3 * i;
        float const norm = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        if (!(norm > 0)) {
            set_error("hx_planar_wall_init: zero surface normal");
            return HX_E_INVALID;
        }
        for (int d = 0; d < 3; ++d) w->surface_normal[i][d] = n[d] / norm;
        w->offset[i] = offset[i];
        for (unsigned j = 0; j < nspecies; ++j) {
            // planar_wall.cpp:96-106: param_potential[j * nwall + i]
            w->epsilon[j * nwall + i] = epsilon[i * nspecies + j];
            w->sigma[j * nwall + i] = sigma[i * nspecies + j];
            w->wetting[j * nwall + i] = wetting[i * nspecies + j];
            w->cutoff[j * nwall + i] = cutoff[i * nspecies + j];
        }
    }
    return HX_OK;
}

int hx_external_planar_wall(const hx_planar_wall* wall, const hx_float4* position_hi_dev, size_t nparticle,
                            const float box_length[3], int force_zero, hx_float4* force_dev, float* en_pot_dev,
                            float* stress_dev, size_t stress_stride, const uint32_t* ctl_dev, void* stream)
{
    if (!wall || !position_hi_dev || !box_length || !force_dev || nparticle >= 0x7FFFFFFFull || wall->nwall == 0
        || wall->nwall > HX_MAX_WALLS || wall->nspecies == 0 || wall->nspecies > HX_MAX_WALL_SPECIES
        || (en_pot_dev == nullptr) != (stress_dev == nullptr)) {
        set_error("hx_external_planar_wall: invalid argument");
        return HX_E_INVALID;
    }
    if (nparticle == 0) return HX_OK;
    wall_args a;
    a.w = *wall;
    a.pos = reinterpret_cast<const float4*>(position_hi_dev);
    a.force = reinterpret_cast<float4*>(force_dev);
    a.en_pot = en_pot_dev;
    a.stress = stress_dev;
    a.ctl = ctl_dev;
    a.stress_stride = stress_stride;
    a.n = static_cast<unsigned>(nparticle);
    a.L = make_vec3(box_length);
    a.force_zero = force_zero;
    unsigned const grid = div_up(nparticle, 128);
    if (en_pot_dev) k_planar_wall<true><<<grid, 128, 0, as_stream(stream)>>>(a);
    else k_planar_wall<false><<<grid, 128, 0, as_stream(stream)>>>(a);
    return check_launch("k_planar_wall");
}

} // extern "C"
