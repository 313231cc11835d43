/*
 * host_backend.c — CPU restatement of HALMD's *host* backend step loop (double precision).
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (see hx_oracle.c header).  Used by bench.py's `cpu_baseline` leg
 * and `--impl reference` arm, and by tests as a physics cross-check (energy conservation, agreement with the
 * GPU-semantics path to rounding).  The reference's own host backend cannot be built here (Boost/Lua/HDF5
 * absent), so this file restates it:
 *   halmd/mdsim/box.hpp:115-133                        minimum image by a single +-L branch
 *   halmd/mdsim/host/binning.cpp:42-116                ncell = floor(L / (r_c,max + skin)), cell lists
 *   halmd/mdsim/host/neighbours/from_binning.cpp:149-229  half lists: 13 half-shell cells + j > i in own cell,
 *                                                      acceptance rr < (r_c + skin)^2
 *   halmd/mdsim/host/max_displacement.cpp:50-86        sqrt(max |r - r0|^2) > skin / 2 triggers a rebuild
 *   halmd/mdsim/host/forces/pair_trunc.hpp:218-344     Newton's third law, aux variables with weight / 2
 *   halmd/mdsim/host/potentials/pair/lennard_jones.hpp:53-63, truncations/{shifted,smooth_r4}.hpp
 *   halmd/mdsim/host/integrators/verlet.cpp:59-103     velocity-Verlet
 * The reference host backend is single-threaded; `nthreads > 1` enables OUR OpenMP parallelisation of the force
 * and neighbour loops (per-thread force buffers), reported as such.
 * Parity unpinned by the reference's tests except for the potential functors (golden tables, see hx_oracle.c).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXT 4

typedef struct host_sim {
    size_t n;
    double L[3], Lh[3];
    unsigned ntypes;
    int trunc;
    double eps[MAXT * MAXT], sigma2[MAXT * MAXT], r_cut[MAXT * MAXT], rr_cut[MAXT * MAXT], en_cut[MAXT * MAXT];
    double rr_cut_skin[MAXT * MAXT];
    double rri_smooth, skin, dt;
    int nthreads;
    double *pos, *vel, *force, *mass, *r0, *en_pot, *virial;
    int* image;
    unsigned* species;
    /* cells */
    unsigned nc[3];
    double cl[3];
    unsigned *cell_start, *cell_items, *cell_of;
    /* half neighbour lists (CSR) */
    size_t* nl_start;
    unsigned* nl_items;
    size_t nl_capacity;
    double* tforce; /* per-thread force buffers */
    int lists_valid;
    size_t rebuilds, steps;
} host_sim;

static inline void potential(const host_sim* s, unsigned idx, double rr, double* fval, double* en)
{
    double sigma2 = s->sigma2[idx];
    double rri = sigma2 / rr;
    double r6i = rri * rri * rri;
    double eps_r6i = s->eps[idx] * r6i;
    double f = 48 * rri * eps_r6i * (r6i - 0.5) / sigma2;
    double e = 4 * eps_r6i * (r6i - 1);
    if (s->trunc == 1) {
        e -= s->en_cut[idx];
    } else if (s->trunc == 2) {
        e -= s->en_cut[idx];
        double r = sqrt(rr);
        double dr = r - s->r_cut[idx];
        double x2 = dr * dr * s->rri_smooth;
        double x4 = x2 * x2;
        double x4i = 1 / (1 + x4);
        double h0 = x4 * x4i;
        double h1 = 4 * dr * s->rri_smooth * x2 * x4i * x4i;
        f = h0 * f - h1 * (e / r);
        e = h0 * e;
    }
    *fval = f;
    *en = e;
}

static inline void reduce_periodic(const host_sim* s, double r[3])
{
    for (int d = 0; d < 3; ++d) {
        if (r[d] > s->Lh[d]) r[d] -= s->L[d];
        else if (r[d] < -s->Lh[d]) r[d] += s->L[d];
    }
}

host_sim* host_sim_create(size_t n, const double L[3], unsigned ntypes, const double* eps, const double* sigma,
                          const double* cutoff_sigma, int trunc, double h, double skin, double dt, int nthreads)
{
    if (ntypes == 0 || ntypes > MAXT) return NULL;
    host_sim* s = (host_sim*)calloc(1, sizeof(host_sim));
    s->n = n;
    s->ntypes = ntypes;
    s->trunc = trunc;
    s->skin = skin;
    s->dt = dt;
    s->rri_smooth = trunc == 2 ? pow(h, -2) : 0;
    double r_cut_max = 0;
    for (unsigned i = 0; i < ntypes * ntypes; ++i) {
        s->eps[i] = eps[i];
        s->sigma2[i] = sigma[i] * sigma[i];
        s->r_cut[i] = sigma[i] * cutoff_sigma[i];
        s->rr_cut[i] = s->r_cut[i] * s->r_cut[i];
        double rri = s->sigma2[i] / s->rr_cut[i], r6i = rri * rri * rri;
        s->en_cut[i] = trunc == 0 ? 0 : 4 * (eps[i] * r6i) * (r6i - 1);
        double rcs = s->r_cut[i] + skin;
        s->rr_cut_skin[i] = rcs * rcs;
        if (rcs > r_cut_max) r_cut_max = rcs;
    }
    size_t ncells = 1;
    for (int d = 0; d < 3; ++d) {
        s->L[d] = L[d];
        s->Lh[d] = L[d] / 2;
        unsigned c = (unsigned)(L[d] / r_cut_max);
        s->nc[d] = c > 1 ? c : 1;
        s->cl[d] = L[d] / s->nc[d];
        ncells *= s->nc[d];
    }
    if (s->nc[0] < 3 || s->nc[1] < 3 || s->nc[2] < 3) { free(s); return NULL; }
#ifdef _OPENMP
    s->nthreads = nthreads > 0 ? nthreads : 1;
#else
    s->nthreads = 1;
#endif
    s->pos = (double*)calloc(3 * n, sizeof(double));
    s->vel = (double*)calloc(3 * n, sizeof(double));
    s->force = (double*)calloc(3 * n, sizeof(double));
    s->r0 = (double*)calloc(3 * n, sizeof(double));
    s->mass = (double*)calloc(n, sizeof(double));
    s->en_pot = (double*)calloc(n, sizeof(double));
    s->virial = (double*)calloc(n, sizeof(double));
    s->image = (int*)calloc(3 * n, sizeof(int));
    s->species = (unsigned*)calloc(n, sizeof(unsigned));
    s->cell_start = (unsigned*)calloc(ncells + 1, sizeof(unsigned));
    s->cell_items = (unsigned*)calloc(n, sizeof(unsigned));
    s->cell_of = (unsigned*)calloc(n, sizeof(unsigned));
    s->nl_start = (size_t*)calloc(n + 1, sizeof(size_t));
    s->nl_capacity = n * 64 + 1024;
    s->nl_items = (unsigned*)malloc(s->nl_capacity * sizeof(unsigned));
    s->tforce = s->nthreads > 1 ? (double*)calloc((size_t)s->nthreads * 3 * n, sizeof(double)) : NULL;
    for (size_t i = 0; i < n; ++i) s->mass[i] = 1;
    return s;
}

void host_sim_destroy(host_sim* s)
{
    if (!s) return;
    free(s->pos); free(s->vel); free(s->force); free(s->r0); free(s->mass); free(s->en_pot); free(s->virial);
    free(s->image); free(s->species); free(s->cell_start); free(s->cell_items); free(s->cell_of);
    free(s->nl_start); free(s->nl_items); free(s->tforce);
    free(s);
}

void host_sim_set(host_sim* s, const double* pos, const double* vel, const unsigned* species, const double* mass)
{
    memcpy(s->pos, pos, 3 * s->n * sizeof(double));
    memcpy(s->vel, vel, 3 * s->n * sizeof(double));
    if (species) memcpy(s->species, species, s->n * sizeof(unsigned));
    if (mass) memcpy(s->mass, mass, s->n * sizeof(double));
    memset(s->image, 0, 3 * s->n * sizeof(int));
    s->lists_valid = 0;
}

void host_sim_get(const host_sim* s, double* pos, double* vel, double* force, int* image)
{
    if (pos) memcpy(pos, s->pos, 3 * s->n * sizeof(double));
    if (vel) memcpy(vel, s->vel, 3 * s->n * sizeof(double));
    if (force) memcpy(force, s->force, 3 * s->n * sizeof(double));
    if (image) memcpy(image, s->image, 3 * s->n * sizeof(int));
}

/* host/binning.cpp:91-116: cell = (uint)(r / l + n) % n */
static void binning_update(host_sim* s)
{
    size_t ncells = (size_t)s->nc[0] * s->nc[1] * s->nc[2];
    memset(s->cell_start, 0, (ncells + 1) * sizeof(unsigned));
    for (size_t i = 0; i < s->n; ++i) {
        unsigned c[3];
        for (int d = 0; d < 3; ++d) c[d] = (unsigned)(s->pos[3 * i + d] / s->cl[d] + (double)s->nc[d]) % s->nc[d];
        unsigned ci = (c[2] * s->nc[1] + c[1]) * s->nc[0] + c[0];
        s->cell_of[i] = ci;
        s->cell_start[ci + 1]++;
    }
    for (size_t c = 0; c < ncells; ++c) s->cell_start[c + 1] += s->cell_start[c];
    unsigned* fill = (unsigned*)calloc(ncells, sizeof(unsigned));
    for (size_t i = 0; i < s->n; ++i) {
        unsigned ci = s->cell_of[i];
        s->cell_items[s->cell_start[ci] + fill[ci]++] = (unsigned)i;   /* push_back order = particle index order */
    }
    free(fill);
}

static size_t neighbours_of(const host_sim* s, size_t i, unsigned* out, size_t cap)
{
    /* host/neighbours/from_binning.cpp:149-229 */
    size_t cnt = 0;
    unsigned ci = s->cell_of[i];
    int cx = (int)(ci % s->nc[0]), cy = (int)((ci / s->nc[0]) % s->nc[1]), cz = (int)(ci / s->nc[0] / s->nc[1]);
    unsigned a = s->species[i];
    for (int k = 0; k <= 13; ++k) {
        int j0 = k / 9 - 1, j1 = (k / 3) % 3 - 1, j2 = k % 3 - 1;   /* k == 13 is (0,0,0) */
        unsigned nx = (unsigned)((cx + (int)s->nc[0] + j0) % (int)s->nc[0]);
        unsigned ny = (unsigned)((cy + (int)s->nc[1] + j1) % (int)s->nc[1]);
        unsigned nz = (unsigned)((cz + (int)s->nc[2] + j2) % (int)s->nc[2]);
        unsigned c2 = (nz * s->nc[1] + ny) * s->nc[0] + nx;
        for (unsigned e = s->cell_start[c2]; e < s->cell_start[c2 + 1]; ++e) {
            unsigned j = s->cell_items[e];
            if (k == 13 && j <= i) continue;
            double r[3] = {s->pos[3 * i] - s->pos[3 * j], s->pos[3 * i + 1] - s->pos[3 * j + 1], s->pos[3 * i + 2] - s->pos[3 * j + 2]};
            reduce_periodic(s, r);
            double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            if (rr >= s->rr_cut_skin[a * s->ntypes + s->species[j]]) continue;
            if (cnt < cap) out[cnt] = j;
            cnt++;
        }
    }
    return cnt;
}

static void neighbour_update(host_sim* s)
{
    binning_update(s);
    size_t n = s->n;
    /* pass 1: counts (parallel), pass 2: fill */
    size_t* cnt = (size_t*)calloc(n + 1, sizeof(size_t));
#pragma omp parallel for schedule(dynamic, 256) num_threads(s->nthreads)
    for (long i = 0; i < (long)n; ++i) cnt[i] = neighbours_of(s, (size_t)i, NULL, 0);
    s->nl_start[0] = 0;
    for (size_t i = 0; i < n; ++i) s->nl_start[i + 1] = s->nl_start[i] + cnt[i];
    free(cnt);
    if (s->nl_start[n] > s->nl_capacity) {
        s->nl_capacity = s->nl_start[n] + s->nl_start[n] / 4;
        free(s->nl_items);
        s->nl_items = (unsigned*)malloc(s->nl_capacity * sizeof(unsigned));
    }
#pragma omp parallel for schedule(dynamic, 256) num_threads(s->nthreads)
    for (long i = 0; i < (long)n; ++i)
        neighbours_of(s, (size_t)i, s->nl_items + s->nl_start[i], s->nl_start[i + 1] - s->nl_start[i]);
    memcpy(s->r0, s->pos, 3 * n * sizeof(double));
    s->lists_valid = 1;
    s->rebuilds++;
}

static double displacement(const host_sim* s)
{
    double rr_max = 0;
    for (size_t i = 0; i < s->n; ++i) {
        double r[3] = {s->pos[3 * i] - s->r0[3 * i], s->pos[3 * i + 1] - s->r0[3 * i + 1], s->pos[3 * i + 2] - s->r0[3 * i + 2]};
        reduce_periodic(s, r);
        double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        if (rr > rr_max) rr_max = rr;
    }
    return sqrt(rr_max);
}

static void force_range(const host_sim* s, size_t begin, size_t end, double* force, double* en_pot, double* virial, int aux)
{
    double const weight = 0.5;
    for (size_t i = begin; i < end; ++i) {
        unsigned a = s->species[i];
        for (size_t e = s->nl_start[i]; e < s->nl_start[i + 1]; ++e) {
            unsigned j = s->nl_items[e];
            double r[3] = {s->pos[3 * i] - s->pos[3 * j], s->pos[3 * i + 1] - s->pos[3 * j + 1], s->pos[3 * i + 2] - s->pos[3 * j + 2]};
            reduce_periodic(s, r);
            double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
            unsigned idx = a * s->ntypes + s->species[j];
            if (!(rr < s->rr_cut[idx])) continue;
            double fval, pot;
            potential(s, idx, rr, &fval, &pot);
            for (int d = 0; d < 3; ++d) {
                force[3 * i + d] += r[d] * fval;
                force[3 * j + d] -= r[d] * fval;
            }
            if (aux) {
                double en = weight * pot, vir = weight * fval * rr;
                en_pot[i] += en; en_pot[j] += en;
                virial[i] += vir; virial[j] += vir;
            }
        }
    }
}

void host_sim_force(host_sim* s, int aux)
{
    if (!s->lists_valid || displacement(s) > s->skin / 2) neighbour_update(s);
    size_t n = s->n;
    memset(s->force, 0, 3 * n * sizeof(double));
    if (aux) {
        memset(s->en_pot, 0, n * sizeof(double));
        memset(s->virial, 0, n * sizeof(double));
    }
    if (s->nthreads <= 1 || aux) {
        force_range(s, 0, n, s->force, s->en_pot, s->virial, aux);
        return;
    }
#ifdef _OPENMP
    int nt = s->nthreads;
#pragma omp parallel num_threads(nt)
    {
        int t = omp_get_thread_num();
        double* f = s->tforce + (size_t)t * 3 * n;
        memset(f, 0, 3 * n * sizeof(double));
#pragma omp for schedule(dynamic, 1)
        for (long b = 0; b < (long)((n + 511) / 512); ++b) {
            size_t lo = (size_t)b * 512, hi = lo + 512 < n ? lo + 512 : n;
            force_range(s, lo, hi, f, NULL, NULL, 0);
        }
#pragma omp for schedule(static)
        for (long k = 0; k < (long)(3 * n); ++k) {
            double acc = 0;
            for (int u = 0; u < nt; ++u) acc += s->tforce[(size_t)u * 3 * n + k];
            s->force[k] = acc;
        }
    }
#endif
}

/* host/integrators/verlet.cpp:59-103 + core::mdstep order: integrate, (force), finalize */
void host_sim_run(host_sim* s, size_t steps)
{
    size_t n = s->n;
    double dt = s->dt, dth = dt / 2;
    if (s->steps == 0 && !s->lists_valid) host_sim_force(s, 0);
    for (size_t st = 0; st < steps; ++st) {
        for (size_t i = 0; i < n; ++i) {
            for (int d = 0; d < 3; ++d) {
                s->vel[3 * i + d] += s->force[3 * i + d] * dth / s->mass[i];
                double r = s->pos[3 * i + d] + s->vel[3 * i + d] * dt;
                if (r > s->Lh[d]) { r -= s->L[d]; s->image[3 * i + d] += 1; }
                else if (r < -s->Lh[d]) { r += s->L[d]; s->image[3 * i + d] -= 1; }
                s->pos[3 * i + d] = r;
            }
        }
        host_sim_force(s, 0);
        for (size_t i = 0; i < n; ++i)
            for (int d = 0; d < 3; ++d) s->vel[3 * i + d] += s->force[3 * i + d] * dth / s->mass[i];
        s->steps++;
    }
}

/* observables/host/thermodynamics.cpp: en_kin, en_pot, virial per particle; out = (E_kin/N, E_pot/N, virial/N, vcm[3]) */
void host_sim_thermodynamics(host_sim* s, double out[6])
{
    host_sim_force(s, 1);
    double ek = 0, ep = 0, vir = 0, mv[3] = {0, 0, 0}, m = 0;
    for (size_t i = 0; i < s->n; ++i) {
        double vv = 0;
        for (int d = 0; d < 3; ++d) { vv += s->vel[3 * i + d] * s->vel[3 * i + d]; mv[d] += s->mass[i] * s->vel[3 * i + d]; }
        ek += 0.5 * s->mass[i] * vv;
        ep += s->en_pot[i];
        vir += s->virial[i];
        m += s->mass[i];
    }
    out[0] = ek / s->n; out[1] = ep / s->n; out[2] = vir / s->n;
    for (int d = 0; d < 3; ++d) out[3 + d] = mv[d] / m;
}

size_t host_sim_rebuilds(const host_sim* s) { return s->rebuilds; }
size_t host_sim_pairs(const host_sim* s) { return s->nl_start[s->n]; }
int host_sim_threads(const host_sim* s) { return s->nthreads; }
