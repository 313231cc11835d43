/*
 * hx_oracle.c — CPU restatement of HALMD's GPU-backend MD step-loop algorithms.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped path (halmd_b200/, libhalmd_b200.so) may call,
 * link or import this file; it is used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline.
 *
 * What it restates: the arithmetic of the reference's CUDA kernels ("GPU semantics": single precision,
 * double-single state, `<=` neighbour acceptance, rint-based minimum image), in the operation order that
 * nvcc 12.9 emits for those kernels under the reference's flags (--fmad=true; checked in SASS/PTX of
 * from_binning_kernel.cu: rr = fma(z,z, fma(x,x, y*y)), r -= image*L as one FMA).  The approximate
 * device operations the reference gets from -prec-div=false / -prec-sqrt=false (MUFU.RCP/RSQ based
 * division and square root) cannot be reproduced bit-for-bit on a CPU; they are replaced by the IEEE
 * operations.  Consequences, stated per function below:
 *   - integer results (cell index, Hilbert key, neighbour sets, rand48 streams, permutations) are exact
 *     restatements except for inputs within 1-2 ulp of a decision boundary of an approximate division;
 *   - floating-point results (forces, energies, positions) agree with the reference kernels to a few ulp.
 * Pinning: the pair-potential functions are checked against the golden tables of the reference's unit tests
 * (tests/golden/lj_potential.json, from test/unit/mdsim/potentials/pair/lennard_jones.cpp:77-211 and
 * test/unit/mdsim/forces/trunc/smooth_r4.cpp:48-142); dsfun, the Hilbert map and the LJ functors are
 * cross-checked against the reference's own headers compiled into oracle/_ref (see oracle/Makefile).
 * Parity UNPINNED by the reference's tests (no golden vectors exist there): Hilbert order, from_binning
 * neighbour sets, max_displacement, rearrange, blocking scheme, MSD, ISF — for these the oracle is pinned
 * against the compiled reference kernels/headers only where oracle/_ref covers them (hilbert map).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction; fmaf() is written where the device
 * code has an FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PH 0xFFFFFFFFu
#define MAXPT 16

typedef struct { float x, y, z, w; } f4;

static inline uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* float → unsigned with the device's saturating semantics (cvt.rzi.u32.f32) */
static inline uint32_t f2u_rz(float f)
{
    if (!(f > 0.f)) return 0;
    if (f >= 4294967296.f) return 0xFFFFFFFFu;
    return (uint32_t)f;
}

/* ---- double-single arithmetic: halmd/numeric/mp/dsfun.hpp:142-207 -------------------------------- */
typedef struct { float hi, lo; } ds;

static inline ds ds_add(ds a, ds b)
{
    float t1 = a.hi + b.hi;
    float e = t1 - a.hi;
    float t2 = ((b.hi - e) + (a.hi - (t1 - e))) + a.lo + b.lo;
    ds c;
    c.hi = t1 + t2;
    c.lo = t2 - (c.hi - t1);
    return c;
}

static inline ds ds_sub(ds a, ds b)
{
    float t1 = a.hi - b.hi;
    float e = t1 - a.hi;
    float t2 = ((-b.hi - e) + (a.hi - (t1 - e))) + a.lo - b.lo;
    ds c;
    c.hi = t1 + t2;
    c.lo = t2 - (c.hi - t1);
    return c;
}

static inline ds ds_mul(ds a, ds b)
{
    float cona = a.hi * 8193.0f;
    float conb = b.hi * 8193.0f;
    float sa1 = cona - (cona - a.hi);
    float sb1 = conb - (conb - b.hi);
    float sa2 = a.hi - sa1;
    float sb2 = b.hi - sb1;
    float c11 = a.hi * b.hi;
    float c21 = (((sa1 * sb1 - c11) + sa1 * sb2) + sa2 * sb1) + sa2 * sb2;
    float c2 = a.hi * b.lo + a.lo * b.hi;
    float t1 = c11 + c2;
    float e = t1 - c11;
    float t2 = ((c2 - e) + (c11 - (t1 - e))) + c21 + a.lo * b.lo;
    ds c;
    c.hi = t1 + t2;
    c.lo = t2 - (c.hi - t1);
    return c;
}

static inline ds mkds(float hi, float lo) { ds r; r.hi = hi; r.lo = lo; return r; }

void oracle_ds_add(const float* a, const float* b, float* c) { ds r = ds_add(mkds(a[0], a[1]), mkds(b[0], b[1])); c[0] = r.hi; c[1] = r.lo; }
void oracle_ds_sub(const float* a, const float* b, float* c) { ds r = ds_sub(mkds(a[0], a[1]), mkds(b[0], b[1])); c[0] = r.hi; c[1] = r.lo; }
void oracle_ds_mul(const float* a, const float* b, float* c) { ds r = ds_mul(mkds(a[0], a[1]), mkds(b[0], b[1])); c[0] = r.hi; c[1] = r.lo; }

/* ---- minimum image: halmd/mdsim/gpu/box_kernel.cuh:37-45 ------------------------------------------
 * image = rint(fdividef(r, L)); r -= image * L  (one FMA after ptxas).  IEEE division here. */
static inline float minimg(float d, float L) { return fmaf(-rintf(d / L), L, d); }
/* inner_prod(r, r) as compiled: fma(z, z, fma(x, x, y*y)) */
static inline float norm2(float x, float y, float z) { return fmaf(z, z, fmaf(x, x, y * y)); }

/* ---- sizing formulas --------------------------------------------------------------------------------- */
size_t oracle_array_size(size_t n)
{
    /* halmd/mdsim/gpu/particle.cpp:69-88 (max_threads_per_block = 1024) */
    size_t a = n > 0 ? (n + 1023) / 1024 : 1;
    return a * 1024;
}

void oracle_binning_plan(const double L[3], size_t n, float r_cut_max, float skin, double occupancy,
                         unsigned ncell[3], float cell_length[3], unsigned* cell_size)
{
    /* halmd/mdsim/gpu/binning.cpp:81-101 */
    double nwarps = n / (occupancy * 32);
    double volume = L[0] * L[1] * L[2];
    size_t ncells = 1;
    for (int d = 0; d < 3; ++d) {
        unsigned ncell_max = (unsigned)(L[d] / (double)(r_cut_max + skin));
        unsigned want = (unsigned)ceil(L[d] * pow(nwarps / volume, 1. / 3));
        ncell[d] = want < ncell_max ? want : ncell_max;
        cell_length[d] = (float)L[d] / (float)ncell[d];
        ncells *= ncell[d];
    }
    *cell_size = 32 * (unsigned)ceil(nwarps / ncells);
}

unsigned oracle_nlist_size(float r_cut_max, float skin, size_t n2, double volume, double occupancy, unsigned cell_size)
{
    /* halmd/mdsim/gpu/neighbours/from_binning.cpp:96-112 */
    float unit_sphere = (float)(4 * M_PI / 3);
    float neighbour_sphere = (float)(unit_sphere * pow(r_cut_max + skin, 3));
    float density = (float)(n2 / volume);
    size_t s = (size_t)ceil(neighbour_sphere * (density / occupancy));
    return (unsigned)(s > cell_size ? s : cell_size);
}

unsigned oracle_hilbert_depth(const double L[3])
{
    /* halmd/mdsim/gpu/sorts/hilbert.cpp:51-54 */
    double m = L[0] > L[1] ? L[0] : L[1];
    if (L[2] > m) m = L[2];
    float max_length = (float)m;
    unsigned depth = (unsigned)ceil(logf(max_length) / M_LN2);
    return depth < 10 ? depth : 10;
}

/* ---- the reference build's division by a constant --------------------------------------------------------------
 * Under -prec-div=false (cmake/platform.cmake:75) r / l is r * MUFU.RCP(l).  MUFU.RCP is a hardware approximation
 * (<= 1 ulp) that cannot be restated on a CPU, but the divisors on this path are a handful of constants (cell
 * lengths, box lengths): GPU tests install a callback that evaluates MUFU.RCP for them on the device (the product's
 * hx_device_rcp); the values are cached here and everything else is restated in C.  Without a callback (CPU-only
 * tests) the IEEE reciprocal stands in, which moves only particles within an ulp of a cell face / lattice plane. */
static struct { float x, rcp; } rcp_table[64];
static unsigned rcp_count = 0;
typedef float (*oracle_rcp_fn)(float);
static oracle_rcp_fn rcp_fn = NULL;

void oracle_set_device_rcp_fn(oracle_rcp_fn fn)
{
    rcp_fn = fn;
    rcp_count = 0;
}

void oracle_set_device_rcp(const float* x, const float* rcp, unsigned count)
{
    for (unsigned k = 0; k < count; ++k) {
        unsigned slot = rcp_count < 64 ? rcp_count : 0;
        for (unsigned q = 0; q < rcp_count; ++q)
            if (fbits(rcp_table[q].x) == fbits(x[k])) slot = q;
        rcp_table[slot].x = x[k];
        rcp_table[slot].rcp = rcp[k];
        if (slot == rcp_count) rcp_count++;
    }
}

static inline float device_rcp(float x)
{
    for (unsigned q = 0; q < rcp_count; ++q)
        if (fbits(rcp_table[q].x) == fbits(x)) return rcp_table[q].rcp;
    if (rcp_fn) {
        float r = rcp_fn(x);
        oracle_set_device_rcp(&x, &r, 1);
        return r;
    }
    return 1.0f / x;
}

/* ---- binning: halmd/mdsim/gpu/binning_kernel.cu:34-139, binning.cpp:166-221 ---------------------------
 * cell = (uint)(r / l + n) % n per dimension, x fastest; as compiled: ONE fused multiply-add rcp(l) * r + n
 * (--fmad=true), then the truncating conversion. */
static inline uint32_t cell_index(f4 p, const float cl[3], const unsigned nc[3])
{
    uint32_t ix = f2u_rz(fmaf(device_rcp(cl[0]), p.x, (float)nc[0])) % nc[0];
    uint32_t iy = f2u_rz(fmaf(device_rcp(cl[1]), p.y, (float)nc[1])) % nc[1];
    uint32_t iz = f2u_rz(fmaf(device_rcp(cl[2]), p.z, (float)nc[2])) % nc[2];
    return (iz * nc[1] + iy) * nc[0] + ix;
}

void oracle_cell_index(const f4* pos, size_t n, const float cl[3], const unsigned nc[3], uint32_t* out)
{
    for (size_t i = 0; i < n; ++i) out[i] = (fbits(pos[i].w) == PH) ? PH : cell_index(pos[i], cl, nc);
}

/* stable sort by cell + assign_cells ⇒ ascending particle index inside each cell, placeholder padding;
 * overflow iff some cell holds >= cell_size particles (last placeholder occupied, binning_kernel.cu:124-127) */
int oracle_binning_update(const f4* pos, size_t n, const float cl[3], const unsigned nc[3], unsigned cell_size,
                          uint32_t* g_cell, uint32_t* cell_count)
{
    size_t ncells = (size_t)nc[0] * nc[1] * nc[2];
    int overflow = 0;
    memset(g_cell, 0xFF, ncells * cell_size * sizeof(uint32_t));
    memset(cell_count, 0, ncells * sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i) {
        if (fbits(pos[i].w) == PH) continue;
        uint32_t c = cell_index(pos[i], cl, nc);
        uint32_t s = cell_count[c]++;
        if (s < cell_size) g_cell[(size_t)c * cell_size + s] = (uint32_t)i;
    }
    for (size_t c = 0; c < ncells; ++c)
        if (cell_count[c] >= cell_size) overflow = 1;
    return overflow;
}

/* ---- Hilbert curve: halmd/mdsim/sorts/hilbert_kernel.hpp:70-142,205-230, gpu/sorts/hilbert_kernel.cu:38-69 -- */
static void tswap(unsigned* v, unsigned* a, unsigned* b)
{
    unsigned va = (*v >> *a) & 7u, vb = (*v >> *b) & 7u;
    *v = *v ^ (va << *a) ^ (vb << *b) ^ (va << *b) ^ (vb << *a);
    unsigned t = *a; *a = *b; *b = t;
}

static uint32_t hilbert_map(float rx, float ry, float rz, unsigned depth)
{
    float r[3] = {rx, ry, rz};
    float n = (float)(1UL << depth);
    for (int d = 0; d < 3; ++d) {
        r[d] = (r[d] - floorf(r[d])) * n;           /* fractional index of the cell in [0, n) */
        r[d] = (floorf(r[d]) + 0.5f) / n;           /* centre of the cell */
        r[d] -= 0.5f;                               /* symmetric coordinates */
    }
    unsigned hcode = 0;
    unsigned a = 21, b = 18, c = 12, d = 15, e = 3, f = 0, g = 6, h = 9;
    unsigned vc = (1U << b) ^ (2U << c) ^ (3U << d) ^ (4U << e) ^ (5U << f) ^ (6U << g) ^ (7U << h);
    for (unsigned i = 0; i < depth; ++i) {
        unsigned x = signbit(r[0]) ? 1u : 0u;
        unsigned y = signbit(r[1]) ? 1u : 0u;
        unsigned z = signbit(r[2]) ? 1u : 0u;
        unsigned v = (vc >> (3 * (x + (y << 1) + (z << 2)))) & 7u;
        r[0] = r[0] * 2 + (float)x - 0.5f;
        r[1] = r[1] * 2 + (float)y - 0.5f;
        r[2] = r[2] * 2 + (float)z - 0.5f;
        if (v == 0) { tswap(&vc, &b, &h); tswap(&vc, &c, &e); }
        else if (v == 1 || v == 2) { tswap(&vc, &c, &g); tswap(&vc, &d, &h); }
        else if (v == 3 || v == 4) { tswap(&vc, &a, &c); tswap(&vc, &f, &h); }
        else if (v == 5 || v == 6) { tswap(&vc, &a, &e); tswap(&vc, &b, &f); }
        else { tswap(&vc, &a, &g); tswap(&vc, &d, &f); }
        hcode = (hcode << 3) + v;
    }
    return hcode;
}

/* unit-box entry point, used to validate against the reference header compiled in oracle/_ref */
uint32_t oracle_hilbert_map_unit(float rx, float ry, float rz, unsigned depth) { return hilbert_map(rx, ry, rz, depth); }

void oracle_hilbert_keys(const f4* pos, size_t nslots, const float L[3], unsigned depth, uint32_t* keys)
{
    for (size_t i = 0; i < nslots; ++i) {
        if (fbits(pos[i].w) == PH) { keys[i] = PH; continue; }
        /* r = element_div(r, box_length): r * rcp(L) as compiled (-prec-div=false), one rounding */
        keys[i] = hilbert_map(pos[i].x * device_rcp(L[0]), pos[i].y * device_rcp(L[1]), pos[i].z * device_rcp(L[2]), depth);
    }
}

/* stable sort of (key, value): halmd/algorithm/gpu/radix_sort.hpp — result equals std::stable_sort by key */
void oracle_sort_pairs_stable(uint32_t* keys, uint32_t* vals, size_t n)
{
    uint32_t* k2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    uint32_t* v2 = (uint32_t*)malloc(n * sizeof(uint32_t));
    for (unsigned pass = 0; pass < 4; ++pass) {
        size_t count[257];
        memset(count, 0, sizeof(count));
        unsigned shift = 8 * pass;
        for (size_t i = 0; i < n; ++i) count[((keys[i] >> shift) & 0xFF) + 1]++;
        for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
        for (size_t i = 0; i < n; ++i) {
            size_t o = count[(keys[i] >> shift) & 0xFF]++;
            k2[o] = keys[i];
            v2[o] = vals[i];
        }
        memcpy(keys, k2, n * sizeof(uint32_t));
        memcpy(vals, v2, n * sizeof(uint32_t));
    }
    free(k2);
    free(v2);
}

/* particle::rearrange: halmd/mdsim/gpu/particle.cpp:173-218, particle_kernel.cu:45-68 */
void oracle_state_permute(const uint32_t* index, size_t n, size_t A, const f4* pos_in, f4* pos_out, const f4* vel_in,
                          f4* vel_out, const f4* img_in, f4* img_out, const uint32_t* id_in, uint32_t* id_out,
                          uint32_t* rid_out)
{
    for (size_t i = 0; i < A; ++i) {
        size_t j = i < n ? index[i] : i;
        pos_out[i] = pos_in[j];
        pos_out[i + A] = pos_in[j + A];
        vel_out[i] = vel_in[j];
        vel_out[i + A] = vel_in[j + A];
        img_out[i] = img_in[j];
        id_out[i] = id_in[j];
    }
    /* iota over the whole array, then values of a stable sort of (id[0..n), iota) */
    uint32_t* keys = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < A; ++i) rid_out[i] = (uint32_t)i;
    memcpy(keys, id_out, n * sizeof(uint32_t));
    oracle_sort_pairs_stable(keys, rid_out, n);
    free(keys);
}

/* max_displacement: halmd/mdsim/gpu/max_displacement_kernel.cu:36-68 → max rr (float) */
float oracle_displacement_max(const f4* pos, const f4* r0, size_t n, const float L[3])
{
    float rr = 0.f;
    for (size_t i = 0; i < n; ++i) {
        float dx = minimg(pos[i].x - r0[i].x, L[0]);
        float dy = minimg(pos[i].y - r0[i].y, L[1]);
        float dz = minimg(pos[i].z - r0[i].z, L[2]);
        float v = norm2(dx, dy, dz);
        if (v > rr) rr = v;
    }
    return rr;
}

/* ---- neighbour lists: halmd/mdsim/gpu/neighbours/from_binning_kernel.cu:99-267 --------------------------- */
static inline unsigned wrapc(int c, int d, int n) { int v = (c + n + d) % n; return (unsigned)v; }

int oracle_nlist_build(const uint32_t* g_cell, const unsigned nc[3], unsigned cell_size, const f4* pos, size_t n,
                       const float* rr_cut_skin, unsigned ntype2, const float L[3], uint32_t* list, unsigned size,
                       size_t stride, int row_layout, uint32_t* count_out)
{
    int overflow = 0;
    size_t ncells = (size_t)nc[0] * nc[1] * nc[2];
    if (row_layout) memset(list, 0xFF, n * stride * sizeof(uint32_t));
    else memset(list, 0xFF, (size_t)size * stride * sizeof(uint32_t));
    for (size_t cell = 0; cell < ncells; ++cell) {
        int cx = (int)(cell % nc[0]), cy = (int)((cell / nc[0]) % nc[1]), cz = (int)(cell / nc[0] / nc[1]);
        for (unsigned s = 0; s < cell_size; ++s) {
            uint32_t p = g_cell[cell * cell_size + s];
            if (p == PH) continue;     /* placeholder threads return after staging */
            f4 r1 = pos[p];
            uint32_t t1 = fbits(r1.w);
            unsigned count = 0;
            /* 13 pairs of opposite cells in loop order j0 (x), j1 (y), j2 (z), then the own cell */
            for (int k = 0; k <= 26; ++k) {
                int dx = 0, dy = 0, dz = 0;
                if (k < 26) {
                    int pair = k >> 1, sg = (k & 1) ? -1 : 1;
                    dx = sg * (pair / 9 - 1);
                    dy = sg * ((pair / 3) % 3 - 1);
                    dz = sg * (pair % 3 - 1);
                }
                size_t c2 = ((size_t)wrapc(cz, dz, (int)nc[2]) * nc[1] + wrapc(cy, dy, (int)nc[1])) * nc[0] + wrapc(cx, dx, (int)nc[0]);
                for (unsigned e = 0; e < cell_size; ++e) {
                    uint32_t m = g_cell[c2 * cell_size + e];
                    if (m == PH) break;
                    if (k == 26 && m == p) continue;
                    f4 r2 = pos[m];
                    float ddx = minimg(r1.x - r2.x, L[0]);
                    float ddy = minimg(r1.y - r2.y, L[1]);
                    float ddz = minimg(r1.z - r2.z, L[2]);
                    float rr = norm2(ddx, ddy, ddz);
                    if (rr <= rr_cut_skin[t1 * ntype2 + fbits(r2.w)] && count < size) {
                        if (row_layout) list[(size_t)p * stride + count] = m;
                        else list[(size_t)count * stride + p] = m;
                        count++;
                    }
                }
            }
            if (count == size) overflow = 1;
            if (count_out) count_out[p] = count;
        }
    }
    return overflow;
}

/* ---- Lennard-Jones + truncations (float = GPU functors, double = host functors) --------------------------- */
typedef struct {
    unsigned ntype1, ntype2;
    int truncation;            /* 0 sharp, 1 shifted, 2 smooth_r4, 3 force_shifted */
    float rri_smooth;
    float epsilon[MAXPT], sigma2[MAXPT], r_cut[MAXPT], rr_cut[MAXPT], en_cut[MAXPT], force_cut[MAXPT];
    int hard_core;             /* adapters/hard_core_kernel.hpp:37-84 */
    float r_core[MAXPT];
} lj_params;

/* lennard_jones_kernel.hpp:43-56 */
static inline void lj_f(float rr, float sigma2, float eps, float* fval, float* en)
{
    float rri = sigma2 / rr;
    float ri6 = rri * rri * rri;
    float eps_ri6 = eps * ri6;
    *fval = 48 * rri * eps_ri6 * (ri6 - 0.5f) / sigma2;
    *en = 4 * eps_ri6 * (ri6 - 1);
}

/* adapters/hard_core_kernel.hpp:66-76 (device) = adapters/hard_core.hpp:86-94 (host, float) */
static inline void lj_core_f(float rr, float r_core, float sigma2, float eps, float* fval, float* en)
{
    float r = sqrtf(rr);
    float r_s = r - r_core;
    lj_f(r_s * r_s, sigma2, eps, fval, en);
    *fval *= r_s / r;
}

void oracle_lj_init_hard_core(lj_params* p, unsigned nt1, unsigned nt2, const float* eps, const float* sigma,
                              const float* core_sigma, int trunc, const float* cutoff_sigma, float h);

void oracle_lj_init(lj_params* p, unsigned nt1, unsigned nt2, const float* eps, const float* sigma, int trunc,
                    const float* cutoff_sigma, float h)
{
    oracle_lj_init_hard_core(p, nt1, nt2, eps, sigma, NULL, trunc, cutoff_sigma, h);
}

void oracle_lj_init_hard_core(lj_params* p, unsigned nt1, unsigned nt2, const float* eps, const float* sigma,
                              const float* core_sigma, int trunc, const float* cutoff_sigma, float h)
{
    /* lennard_jones.cpp:52-70, smooth_r4.hpp:55-91 */
    memset(p, 0, sizeof(*p));
    p->ntype1 = nt1;
    p->ntype2 = nt2;
    p->truncation = trunc;
    p->rri_smooth = trunc == 2 ? (float)pow((double)h, -2) : 0.f;
    p->hard_core = core_sigma != NULL;
    for (unsigned i = 0; i < nt1 * nt2; ++i) {
        p->epsilon[i] = eps[i];
        p->sigma2[i] = sigma[i] * sigma[i];
        p->r_cut[i] = sigma[i] * cutoff_sigma[i];
        p->rr_cut[i] = p->r_cut[i] * p->r_cut[i];
        float f, e;
        if (core_sigma) {
            p->r_core[i] = core_sigma[i] * sigma[i];       /* element_prod(core, sigma), adapters/hard_core.hpp:60 */
            lj_core_f(p->rr_cut[i], p->r_core[i], p->sigma2[i], p->epsilon[i], &f, &e);
        } else {
            lj_f(p->rr_cut[i], p->sigma2[i], p->epsilon[i], &f, &e);
        }
        p->en_cut[i] = trunc == 0 ? 0.f : e;
        p->force_cut[i] = trunc == 3 ? f * p->r_cut[i] : 0.f;      /* force_shifted.hpp:69-70 */
    }
}

/* potential(rr) with truncation; returns 0 outside the cutoff (within_range is rr < rr_cut) */
int oracle_pair_f(const lj_params* p, unsigned idx, float rr, float* fval, float* en)
{
    if (!(rr < p->rr_cut[idx])) return 0;
    if (p->hard_core) lj_core_f(rr, p->r_core[idx], p->sigma2[idx], p->epsilon[idx], fval, en);
    else lj_f(rr, p->sigma2[idx], p->epsilon[idx], fval, en);
    if (p->truncation == 1) {
        *en = *en - p->en_cut[idx];                    /* shifted_kernel.cuh:49-57 */
    } else if (p->truncation == 2) {
        /* smooth_r4_kernel.cuh:53-76 */
        float e = *en - p->en_cut[idx];
        float r = sqrtf(rr);
        float dr = r - p->r_cut[idx];
        float x2 = dr * dr * p->rri_smooth;
        float x4 = x2 * x2;
        float x4i = 1 / (1 + x4);
        float h0 = x4 * x4i;
        float h1 = 4 * dr * p->rri_smooth * x2 * x4i * x4i;
        *fval = h0 * *fval - h1 * (e / r);
        *en = h0 * e;
    } else if (p->truncation == 3) {
        /* force_shifted_kernel.cuh:52-63 */
        float r = sqrtf(rr);
        *fval = *fval - p->force_cut[idx] / r;
        *en = *en - p->en_cut[idx] + (r - p->r_cut[idx]) * p->force_cut[idx];
    }
    return 1;
}

/* host functors in double: halmd/mdsim/host/potentials/pair/lennard_jones.hpp:53-63,
 * truncations/{sharp,shifted,smooth_r4}.hpp — the reference's golden tables are for these */
void oracle_pair_d(double eps, double sigma, double cutoff_sigma, int trunc, double h, double rr, double* fval, double* en)
{
    double sigma2 = sigma * sigma;
    double r_cut = sigma * cutoff_sigma, rr_cut = r_cut * r_cut;
    double rri = sigma2 / rr_cut, r6i = rri * rri * rri;
    double en_cut = 4 * (eps * r6i) * (r6i - 1);
    rri = sigma2 / rr;
    r6i = rri * rri * rri;
    double eps_r6i = eps * r6i;
    double f = 48 * rri * eps_r6i * (r6i - 0.5) / sigma2;
    double e = 4 * eps_r6i * (r6i - 1);
    if (trunc == 1) {
        e -= en_cut;
    } else if (trunc == 2) {
        double rri_smooth = pow(h, -2);
        e -= en_cut;
        double r = sqrt(rr);
        double dr = r - r_cut;
        double x2 = dr * dr * rri_smooth;
        double x4 = x2 * x2;
        double x4i = 1 / (1 + x4);
        double h0 = x4 * x4i;
        double h1 = 4 * dr * rri_smooth * x2 * x4i * x4i;
        f = h0 * f - h1 * (e / r);
        e = h0 * e;
    } else if (trunc == 3) {
        /* host/potentials/pair/truncations/force_shifted.hpp: F_c = fval(r_c) r_c */
        double rric = sigma2 / rr_cut, r6ic = rric * rric * rric;
        double force_cut = 48 * rric * (eps * r6ic) * (r6ic - 0.5) / sigma2 * r_cut;
        double r = sqrt(rr);
        f -= force_cut / r;
        e = e - en_cut + (r - r_cut) * force_cut;
    }
    *fval = f;
    *en = e;
}

/* ---- pair force: halmd/mdsim/gpu/forces/pair_trunc_kernel.cuh:159-255 ---------------------------------------
 * sequential walk of the list until the first placeholder, force accumulated in double-single, aux in float */
void oracle_pair_force(const lj_params* p, const f4* pos1, const f4* pos2, const uint32_t* list, unsigned size,
                       size_t stride, int row_layout, size_t n, const float L[3], int force_zero, float aux_weight,
                       f4* force, float* en_pot, float* stress, size_t stress_stride)
{
    for (size_t i = 0; i < n; ++i) {
        f4 r1 = pos1[i];
        uint32_t t1 = fbits(r1.w);
        ds fx = mkds(0, 0), fy = mkds(0, 0), fz = mkds(0, 0);
        float en_sum = 0;
        float s[6] = {0, 0, 0, 0, 0, 0};
        for (unsigned k = 0; k < size; ++k) {
            uint32_t j = row_layout ? list[i * stride + k] : list[(size_t)k * stride + i];
            if (j == PH) break;
            f4 r2 = pos2[j];
            float dx = minimg(r1.x - r2.x, L[0]);
            float dy = minimg(r1.y - r2.y, L[1]);
            float dz = minimg(r1.z - r2.z, L[2]);
            float rr = norm2(dx, dy, dz);
            float fval, en;
            if (!oracle_pair_f(p, t1 * p->ntype2 + fbits(r2.w), rr, &fval, &en)) continue;
            fx = ds_add(fx, mkds(fval * dx, 0));
            fy = ds_add(fy, mkds(fval * dy, 0));
            fz = ds_add(fz, mkds(fval * dz, 0));
            if (en_pot) {
                en_sum += aux_weight * en;
                float w = aux_weight * fval;
                s[0] += w * (dx * dx);
                s[1] += w * (dy * dy);
                s[2] += w * (dz * dz);
                s[3] += w * (dx * dy);
                s[4] += w * (dx * dz);
                s[5] += w * (dy * dz);
            }
        }
        if (!force_zero) {
            fx = ds_add(fx, mkds(force[i].x, 0));
            fy = ds_add(fy, mkds(force[i].y, 0));
            fz = ds_add(fz, mkds(force[i].z, 0));
            if (en_pot) {
                en_sum += en_pot[i];
                for (int k = 0; k < 6; ++k) s[k] += stress[k * stress_stride + i];
            }
        }
        force[i].x = fx.hi;
        force[i].y = fy.hi;
        force[i].z = fz.hi;
        force[i].w = 0;
        if (en_pot) {
            en_pot[i] = en_sum;
            for (int k = 0; k < 6; ++k) stress[k * stress_stride + i] = s[k];
        }
    }
}

/* ---- velocity-Verlet: halmd/mdsim/gpu/integrators/verlet_kernel.cu:37-98 ------------------------------------ */
static inline ds half_kick(ds v, float f, float half_dt, float mass) { return ds_add(v, mkds(f * half_dt / mass, 0)); }

void oracle_verlet_finalize(f4* vel, const f4* force, size_t A, float dt)
{
    float half_dt = dt / 2;
    for (size_t i = 0; i < A; ++i) {
        float mass = vel[i].w;
        ds vx = half_kick(mkds(vel[i].x, vel[i + A].x), force[i].x, half_dt, mass);
        ds vy = half_kick(mkds(vel[i].y, vel[i + A].y), force[i].y, half_dt, mass);
        ds vz = half_kick(mkds(vel[i].z, vel[i + A].z), force[i].z, half_dt, mass);
        vel[i].x = vx.hi; vel[i].y = vy.hi; vel[i].z = vz.hi;
        vel[i + A].x = vx.lo; vel[i + A].y = vy.lo; vel[i + A].z = vz.lo; vel[i + A].w = 0;
    }
}

static inline float drift_wrap(ds* r, ds v, float dt, float L)
{
    *r = ds_add(*r, ds_mul(v, mkds(dt, 0)));
    float image = rintf(r->hi / L);                 /* rint(fdividef(r, L)) with IEEE division */
    *r = ds_sub(*r, mkds(image * L, 0));
    return image;
}

void oracle_verlet_integrate(f4* pos, f4* image, f4* vel, const f4* force, size_t A, float dt, const float L[3])
{
    float half_dt = dt / 2;
    for (size_t i = 0; i < A; ++i) {
        float mass = vel[i].w;
        ds vx = half_kick(mkds(vel[i].x, vel[i + A].x), force[i].x, half_dt, mass);
        ds vy = half_kick(mkds(vel[i].y, vel[i + A].y), force[i].y, half_dt, mass);
        ds vz = half_kick(mkds(vel[i].z, vel[i + A].z), force[i].z, half_dt, mass);
        ds rx = mkds(pos[i].x, pos[i + A].x), ry = mkds(pos[i].y, pos[i + A].y), rz = mkds(pos[i].z, pos[i + A].z);
        float ix = drift_wrap(&rx, vx, dt, L[0]);
        float iy = drift_wrap(&ry, vy, dt, L[1]);
        float iz = drift_wrap(&rz, vz, dt, L[2]);
        pos[i].x = rx.hi; pos[i].y = ry.hi; pos[i].z = rz.hi;
        pos[i + A].x = rx.lo; pos[i + A].y = ry.lo; pos[i + A].z = rz.lo; pos[i + A].w = 0;
        vel[i].x = vx.hi; vel[i].y = vy.hi; vel[i].z = vz.hi;
        vel[i + A].x = vx.lo; vel[i + A].y = vy.lo; vel[i + A].z = vz.lo; vel[i + A].w = 0;
        if (!(ix == 0 && iy == 0 && iz == 0)) {
            image[i].x = ix + image[i].x;
            image[i].y = iy + image[i].y;
            image[i].z = iz + image[i].z;
            image[i].w = 0;
        }
    }
}

/* ---- Nose-Hoover chain: halmd/mdsim/gpu/integrators/verlet_nvt_hoover_kernel.cu:37-125 (kernels),
 * verlet_nvt_hoover.cpp:79-95, :175-222 (host-side chain) --------------------------------------------------------- */

/* v *= scale for dsfloat operands as the reference build executes dsmul (dsfun.hpp:178-207 under --fmad=true): the
 * cross term a0 * b1 + a1 * b0 and the low-order product ride on fused multiply-adds; the Dekker part is exact either
 * way, so only these two contractions are visible in the result */
static inline ds ds_mul_gpu(ds a, ds b)
{
    float cona = a.hi * 8193.0f;
    float conb = b.hi * 8193.0f;
    float sa1 = cona - (cona - a.hi);
    float sb1 = conb - (conb - b.hi);
    float sa2 = a.hi - sa1;
    float sb2 = b.hi - sb1;
    float c11 = a.hi * b.hi;
    float c21 = (((sa1 * sb1 - c11) + sa1 * sb2) + sa2 * sb1) + sa2 * sb2;
    float c2 = fmaf(a.hi, b.lo, a.lo * b.hi);
    float t1 = c11 + c2;
    float e = t1 - c11;
    float t2 = fmaf(a.lo, b.lo, c21 + ((c2 - e) + (c11 - (t1 - e))));
    ds c;
    c.hi = t1 + t2;
    c.lo = t2 - (c.hi - t1);
    return c;
}

/* dsfloat(double): dsfloat.hpp — hi = float(a), lo = float(a - hi) */
static inline ds ds_from_double(double a)
{
    ds r;
    r.hi = (float)a;
    r.lo = (float)(a - (double)r.hi);
    return r;
}

void oracle_hoover_rescale(f4* vel, size_t A, double scale)
{
    ds s = ds_from_double(scale);
    for (size_t i = 0; i < A; ++i) {
        ds vx = ds_mul_gpu(mkds(vel[i].x, vel[i + A].x), s);
        ds vy = ds_mul_gpu(mkds(vel[i].y, vel[i + A].y), s);
        ds vz = ds_mul_gpu(mkds(vel[i].z, vel[i + A].z), s);
        vel[i].x = vx.hi; vel[i].y = vy.hi; vel[i].z = vz.hi;
        vel[i + A].x = vx.lo; vel[i + A].y = vy.lo; vel[i + A].z = vz.lo; vel[i + A].w = 0;
    }
}

void oracle_hoover_integrate(f4* pos, f4* image, f4* vel, const f4* force, size_t A, float dt, double scale, const float L[3])
{
    oracle_hoover_rescale(vel, A, scale);
    oracle_verlet_integrate(pos, image, vel, force, A, dt, L);
}

typedef struct { double xi[2], v_xi[2], mass_xi[2], temperature, en_kin_target_2, dof; } hoover_chain;

void oracle_hoover_chain_init(hoover_chain* c, size_t nparticle, unsigned dimension, double temperature, double resonance_frequency)
{
    memset(c, 0, sizeof(*c));
    c->dof = (double)dimension * (double)nparticle;
    c->temperature = temperature;
    c->en_kin_target_2 = c->dof * temperature;
    double omega_sq = pow(2 * M_PI * resonance_frequency, 2);
    c->mass_xi[0] = c->dof * temperature / omega_sq;
    c->mass_xi[1] = temperature / omega_sq;
}

double oracle_hoover_chain_propagate(hoover_chain* c, double en_kin_2, double timestep)
{
    double timestep_half = timestep / 2, timestep_4 = timestep / 4, timestep_8 = timestep / 8;
    c->v_xi[1] += (c->mass_xi[0] * c->v_xi[0] * c->v_xi[0] - c->temperature) / c->mass_xi[1] * timestep_4;
    double t = exp(-c->v_xi[1] * timestep_8);
    c->v_xi[0] *= t;
    c->v_xi[0] += (en_kin_2 - c->en_kin_target_2) / c->mass_xi[0] * timestep_4;
    c->v_xi[0] *= t;
    for (int i = 0; i < 2; ++i) c->xi[i] += c->v_xi[i] * timestep_half;
    double s = exp(-c->v_xi[0] * timestep_half);
    en_kin_2 *= s * s;
    c->v_xi[0] *= t;
    c->v_xi[0] += (en_kin_2 - c->en_kin_target_2) / c->mass_xi[0] * timestep_4;
    c->v_xi[0] *= t;
    c->v_xi[1] += (c->mass_xi[0] * c->v_xi[0] * c->v_xi[0] - c->temperature) / c->mass_xi[1] * timestep_4;
    return s;
}

double oracle_hoover_chain_energy(const hoover_chain* c, size_t nparticle)
{
    double en = c->temperature * (c->dof * c->xi[0] + c->xi[1]);
    for (int i = 0; i < 2; ++i) en += c->mass_xi[i] * c->v_xi[i] * c->v_xi[i] / 2;
    return en / (double)nparticle;
}

/* ---- rand48: halmd/random/gpu/rand48_kernel.cu:42-100, rand48_kernel.cuh:40-79, uint48.hpp:72-111 ------------ */
#define R48_MASK ((1ull << 48) - 1)
#define R48_A 0x5DEECE66Dull
#define R48_C 0xBull

typedef struct { unsigned blocks, threads; uint16_t a[3], c[3]; } rand48_t;

static inline uint64_t pack48(const uint16_t* s) { return (uint64_t)s[0] | ((uint64_t)s[1] << 16) | ((uint64_t)s[2] << 32); }
static inline void unpack48(uint64_t v, uint16_t* s) { s[0] = v & 0xFFFF; s[1] = (v >> 16) & 0xFFFF; s[2] = (v >> 32) & 0xFFFF; }

/* limb arithmetic exactly as uint48.hpp:72-111 (used to cross-check the 64-bit shortcut) */
void oracle_muladd48_limbs(const uint16_t a[3], const uint16_t b[3], const uint16_t c[3], uint16_t r[3])
{
    unsigned d = (unsigned)a[0] * b[0] + c[0];
    r[0] = d & 0xFFFF;
    d >>= 16;
    d += (unsigned)a[0] * b[1] + (unsigned)a[1] * b[0] + c[1];
    r[1] = d & 0xFFFF;
    d >>= 16;
    d += (unsigned)a[0] * b[2] + (unsigned)a[1] * b[1] + (unsigned)a[2] * b[0] + c[2];
    r[2] = d & 0xFFFF;
}

void oracle_rand48_seed(rand48_t* rng, unsigned blocks, unsigned threads, unsigned seed, uint16_t* state)
{
    unsigned T = blocks * threads;
    rng->blocks = blocks;
    rng->threads = threads;
    uint16_t x[3] = {0x330E, 0xABCD, 0x1234};
    if (seed > 0) { x[1] = seed & 0xFFFF; x[2] = (seed >> 16) & 0xFFFF; }
    /* sequential definition: thread t holds x_{t+1} of the plain LCG; leapfrog constants map x -> x_{+T} */
    uint64_t s = pack48(x);
    uint64_t A = 1, C = 0;
    for (unsigned t = 0; t < T; ++t) {
        s = (R48_A * s + R48_C) & R48_MASK;
        unpack48(s, state + 3 * t);
        A = (R48_A * A) & R48_MASK;
        C = (R48_A * C + R48_C) & R48_MASK;
    }
    unpack48(A, rng->a);
    unpack48(C, rng->c);
}

static inline float r48_uniform(uint64_t* s, uint64_t a, uint64_t c)
{
    float u = (float)((*s >> 32) & 0xFFFF) / 65536.f + (float)((*s >> 16) & 0xFFFF) / 4294967296.f;
    *s = (a * *s + c) & R48_MASK;
    return u;
}

void oracle_rand48_uniform(const rand48_t* rng, uint16_t* state, float* out, size_t n)
{
    unsigned T = rng->blocks * rng->threads;
    uint64_t a = pack48(rng->a), c = pack48(rng->c);
    for (unsigned t = 0; t < T; ++t) {
        uint64_t s = pack48(state + 3 * t);
        for (size_t i = t; i < n; i += T) out[i] = r48_uniform(&s, a, c);
        unpack48(s, state + 3 * t);
    }
}

/* normal_distribution.cuh:49-67 */
static inline void normal_pair(uint64_t* s, uint64_t a, uint64_t c, float sigma, float* n1, float* n2)
{
    float v1, v2, q;
    do {
        v1 = fmaf(2.f, r48_uniform(s, a, c), -1.f);
        v2 = fmaf(2.f, r48_uniform(s, a, c), -1.f);
        q = fmaf(v1, v1, v2 * v2);
    } while (q >= 1.f);
    float w = sigma * sqrtf(-2 * logf(q) / q);
    *n1 = w * v1;
    *n2 = w * v2;
}

/* verlet_nvt_andersen_kernel.cu:96-171, emulating the 32-lane vote of each launch warp.
 * q = 1 - pow(1 - coll_prob, 1/32) evaluated in IEEE float (device: __powf approximation). */
void oracle_andersen_finalize(f4* vel, const f4* force, size_t n, size_t A, float dt, float sqrt_temperature,
                              float coll_prob, const rand48_t* rng, uint16_t* state)
{
    unsigned T = rng->blocks * rng->threads;
    uint64_t a = pack48(rng->a), c = pack48(rng->c);
    float q = 1.f - powf(1.f - coll_prob, 1.f / 32);
    float half_dt = dt / 2;
    for (unsigned w0 = 0; w0 < T; w0 += 32) {
        uint64_t s[32];
        int cached[32];
        float cache[32];
        for (int l = 0; l < 32; ++l) { s[l] = pack48(state + 3 * (w0 + l)); cached[l] = 0; cache[l] = 0; }
        for (size_t base = w0; base < n; base += T) {
            int all = 1;
            float u[32];
            for (int l = 0; l < 32; ++l) {
                size_t i = base + l;
                if (i >= n) continue;
                u[l] = r48_uniform(&s[l], a, c);
                if (!(u[l] > q)) all = 0;
            }
            for (int l = 0; l < 32; ++l) {
                size_t i = base + l;
                if (i >= n) continue;
                float mass = vel[i].w;
                if (all) {
                    ds vx = half_kick(mkds(vel[i].x, vel[i + A].x), force[i].x, half_dt, mass);
                    ds vy = half_kick(mkds(vel[i].y, vel[i + A].y), force[i].y, half_dt, mass);
                    ds vz = half_kick(mkds(vel[i].z, vel[i + A].z), force[i].z, half_dt, mass);
                    vel[i].x = vx.hi; vel[i].y = vy.hi; vel[i].z = vz.hi;
                    vel[i + A].x = vx.lo; vel[i + A].y = vy.lo; vel[i + A].z = vz.lo; vel[i + A].w = 0;
                } else {
                    float vx, vy, vz;
                    normal_pair(&s[l], a, c, sqrt_temperature, &vx, &vy);
                    cached[l] = !cached[l];
                    if (cached[l]) normal_pair(&s[l], a, c, sqrt_temperature, &vz, &cache[l]);
                    else vz = cache[l];
                    vel[i].x = vx; vel[i].y = vy; vel[i].z = vz;
                    vel[i + A].x = 0; vel[i + A].y = 0; vel[i + A].z = 0; vel[i + A].w = 0;
                }
            }
        }
        for (int l = 0; l < 32; ++l) unpack48(s[l], state + 3 * (w0 + l));
    }
}

/* ---- observables ------------------------------------------------------------------------------------------------ */
/* phase_space_kernel.cu:39-66 */
void oracle_sample_positions(const f4* pos, const f4* image, const uint32_t* rid, size_t n, const float L[3], f4* out)
{
    for (size_t i = 0; i < n; ++i) {
        uint32_t r = rid[i];
        out[i].x = fmaf(image[r].x, L[0], pos[r].x);
        out[i].y = fmaf(image[r].y, L[1], pos[r].y);
        out[i].z = fmaf(image[r].z, L[2], pos[r].z);
        out[i].w = pos[r].w;
    }
}

/* mean_square_displacement.hpp:35-41 + accumulator.hpp:56-100, accumulated sequentially in double:
 * out = (count, mean, M2) */
void oracle_msd(const f4* first, const f4* second, size_t n, double out[3])
{
    double cnt = 0, m = 0, v = 0;
    for (size_t i = 0; i < n; ++i) {
        float dx = second[i].x - first[i].x, dy = second[i].y - first[i].y, dz = second[i].z - first[i].z;
        double x = (double)norm2(dx, dy, dz);
        double t = x - m;
        cnt += 1;
        m += t / cnt;
        v += t * (x - m);
    }
    out[0] = cnt; out[1] = m; out[2] = v;
}

/* the other tagged-particle correlation functions, same accumulation: kind 1 mean_quartic_displacement.hpp:36-44,
 * kind 2 velocity_autocorrelation.hpp:36-39 (kind 0 = oracle_msd) */
void oracle_correlate(int kind, const f4* first, const f4* second, size_t n, double out[3])
{
    double cnt = 0, m = 0, v = 0;
    for (size_t i = 0; i < n; ++i) {
        float val;
        if (kind == 2) {
            val = fmaf(first[i].z, second[i].z, fmaf(first[i].x, second[i].x, first[i].y * second[i].y));
        } else {
            float dx = second[i].x - first[i].x, dy = second[i].y - first[i].y, dz = second[i].z - first[i].z;
            float rr = norm2(dx, dy, dz);
            val = kind == 1 ? rr * rr : rr;
        }
        double x = (double)val;
        double t = x - m;
        cnt += 1;
        m += t / cnt;
        v += t * (x - m);
    }
    out[0] = cnt; out[1] = m; out[2] = v;
}

/* density_mode_kernel.cu:46-80 + density_mode.cpp:126-129: (sum cos, -sum sin), summed in double here */
void oracle_density_modes(const f4* pos, size_t n, const float* q, unsigned nq, double* rho)
{
    for (unsigned k = 0; k < nq; ++k) {
        double c = 0, s = 0;
        for (size_t i = 0; i < n; ++i) {
            float qr = fmaf(q[3 * k + 2], pos[i].z, fmaf(q[3 * k], pos[i].x, q[3 * k + 1] * pos[i].y));
            c += (double)cosf(qr);
            s += (double)sinf(qr);
        }
        rho[2 * k] = c;
        rho[2 * k + 1] = -s;
    }
}

/* thermodynamics_kernel.cu:35-97: sums of m v^2, m v, m, en_pot, virial trace */
void oracle_thermodynamics(const f4* vel, const float* en_pot, const float* stress, size_t stress_stride, size_t n,
                           double out[8])
{
    for (int k = 0; k < 8; ++k) out[k] = 0;
    for (size_t i = 0; i < n; ++i) {
        /* thermodynamics_kernel.cu:35-97 as built (--fmad=true): inner_prod(v, v) in float, the product with the mass
         * fused into the double-single accumulation, i.e. unrounded */
        out[0] += (double)vel[i].w * (double)norm2(vel[i].x, vel[i].y, vel[i].z);
        out[1] += (double)vel[i].w * (double)vel[i].x;
        out[2] += (double)vel[i].w * (double)vel[i].y;
        out[3] += (double)vel[i].w * (double)vel[i].z;
        out[4] += (double)vel[i].w;
        if (en_pot) out[5] += (double)en_pot[i];
        if (stress) out[6] += (double)stress[i] + (double)stress[stress_stride + i] + (double)stress[2 * stress_stride + i];
    }
}

/* ---- external one-body force, planar walls --------------------------------------------------------------------
 * halmd/mdsim/gpu/forces/external_kernel.cuh:49-103 + potentials/external/planar_wall_kernel.cu:43-105 in float
 * with IEEE division (the device build divides approximately).  Tables indexed [species * nwall + wall]. */
typedef struct {
    unsigned nwall, nspecies;
    float smoothing;
    float surface_normal[8][3];
    float offset[8];
    float epsilon[32], sigma[32], wetting[32], cutoff[32];
} planar_wall_t;

void oracle_planar_wall(const planar_wall_t* w, const f4* pos, size_t n, const float L[3], int force_zero, f4* force,
                        float* en_pot_out, float* stress, size_t stress_stride)
{
    for (size_t i = 0; i < n; ++i) {
        float r[3] = {pos[i].x, pos[i].y, pos[i].z};
        unsigned species;
        memcpy(&species, &pos[i].w, 4);
        for (int d = 0; d < 3; ++d) r[d] = fmaf(-rintf(r[d] / L[d]), L[d], r[d]);      /* reduce_periodic */
        float en_pot = 0, f[3] = {0, 0, 0};
        for (unsigned k = 0; k < w->nwall; ++k) {
            const float* nrm = w->surface_normal[k];
            /* inner_prod in the order nvcc emits for the reference kernel: y*ny, then x, then z (fused) */
            float d = fmaf(r[2], nrm[2], fmaf(r[0], nrm[0], r[1] * nrm[1])) - w->offset[k];
            unsigned t = species * w->nwall + k;
            float epsilon = w->epsilon[t], sigma = w->sigma[t], wet = w->wetting[t], cutoff = w->cutoff[t];
            if (fabsf(d) >= cutoff) continue;
            float sc = sigma / cutoff;
            float dc3i = sc * sc * sc;
            float dc6i = (2.f / 15) * dc3i * dc3i;
            float en_cut = epsilon * dc3i * (dc6i - wet);
            float sd = sigma / d;
            float d3i = sd * sd * sd;
            float d6i = (2.f / 15) * d3i * d3i;
            float eps_d3i = (d > 0 ? 1.f : -1.f) * epsilon * d3i;
            float en_sub = eps_d3i * (d6i - wet) - en_cut;
            float fval = 3.f * eps_d3i * (3.f * d6i - wet) / d;
            float dd = (fabsf(d) - cutoff) / w->smoothing;
            float x2 = dd * dd, x4 = x2 * x2;
            float x4i = 1.f / (1.f + x4);
            float h0 = x4 * x4i;
            float h1 = 4.f * dd * x2 * x4i * x4i;
            fval = h0 * fval - h1 * en_sub / w->smoothing;
            en_sub = h0 * en_sub;
            for (int c = 0; c < 3; ++c) f[c] = fmaf(fval, nrm[c], f[c]);
            en_pot += en_sub;
        }
        if (!force_zero) {
            f[0] += force[i].x; f[1] += force[i].y; f[2] += force[i].z;
        }
        force[i].x = f[0]; force[i].y = f[1]; force[i].z = f[2]; force[i].w = 0;
        if (en_pot_out) {
            if (!force_zero) en_pot += en_pot_out[i];
            else for (int c = 0; c < 6; ++c) stress[c * stress_stride + i] = 0;
            en_pot_out[i] = en_pot;
        }
    }
}
