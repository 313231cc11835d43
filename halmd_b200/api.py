"""Host-side mirror of HALMD's Lua module API for the MD step loop.

Each class corresponds to the Lua module of the same name (``lua/halmd/**``) and takes the same keyword
arguments; underneath it owns a handle of the C++ module layer (``include/halmd_b200_modules.h``).  A HALMD
script such as ``examples/liquid/lennard_jones.lua`` translates line by line::

    box = mdsim.box(length=[L, L, L])
    particle = mdsim.particle(dimension=3, particles=N)
    potential = mdsim.potentials.pair.lennard_jones().truncate("smooth_r4", cutoff=2.5, h=0.005)
    force = mdsim.forces.pair(box=box, particle=particle, potential=potential)
    integrator = mdsim.integrators.verlet(box=box, particle=particle, timestep=0.001)
    observables.sampler.run(steps)

There is no CPU fallback: constructing ``particle`` without a CUDA device raises :class:`HxError`.
"""
import ctypes as C

import numpy as np

from . import _lib as hx
from ._lib import HxError, check, lib

__all__ = ["mdsim", "observables", "HxError"]


def _matrix(value, n1, n2, name):
    a = np.asarray(value, dtype=np.float32)
    if a.ndim == 0:
        a = np.full((n1, n2), float(a), dtype=np.float32)
    if a.shape != (n1, n2):
        raise ValueError(f"bad argument '{name}': expected a {n1}x{n2} matrix")
    return np.ascontiguousarray(a)


class _Handle:
    _destroy = None

    def __init__(self):
        self._h = C.c_void_p()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and self._destroy:
            try:
                getattr(lib, self._destroy)(h)
            except Exception:
                pass


# ---- mdsim.box (lua/halmd/mdsim/box.lua.in) ------------------------------------------------------------------
class box:
    def __init__(self, length=None, edges=None):
        if (length is None) == (edges is None):
            raise ValueError("box: exactly one of 'length' or 'edges' is required")
        if edges is not None:
            e = np.asarray(edges, dtype=np.float64)
            if e.shape != (3, 3) or np.any(e - np.diag(np.diag(e)) != 0):
                raise ValueError("box: non-cuboid geometry is not implemented")   # mdsim/box.cpp:41-52
            length = np.diag(e)
        self.length = np.asarray(length, dtype=np.float64)
        if self.length.shape != (3,) or np.any(self.length <= 0):
            raise ValueError("box: 'length' must hold 3 positive edge lengths")

    @property
    def volume(self):
        return float(np.prod(self.length))

    @property
    def edges(self):
        return np.diag(self.length)

    def _d3(self):
        return hx.d3(self.length)


# ---- mdsim.particle (lua/halmd/mdsim/particle.lua.in) -----------------------------------------------------------
class particle(_Handle):
    _destroy = "hxm_particle_destroy"
    _arrays = {
        "position": (np.float32, 3), "velocity": (np.float32, 3), "image": (np.float32, 3), "force": (np.float32, 3),
        "species": (np.uint32, 0), "mass": (np.float32, 0), "id": (np.uint32, 0), "reverse_id": (np.uint32, 0),
        "potential_energy": (np.float32, 0), "potential_stress_tensor": (np.float32, 6),
        "position_ds": (np.float64, 3), "velocity_ds": (np.float64, 3),
    }

    def __init__(self, dimension=3, particles=None, species=1, memory="gpu", precision="double-single", stream=None):
        super().__init__()
        if dimension != 3:
            raise ValueError("particle: only dimension=3 is built")
        if memory != "gpu":
            raise HxError("particle: memory='host' is not available (no CPU fallback)")
        if precision not in ("double-single", "dsfloat"):
            raise ValueError("particle: only precision='double-single' is built")
        if not particles or particles < 1:
            raise ValueError("bad argument 'particles'")
        self.dimension, self.memory, self.precision = dimension, memory, precision
        check(lib.hxm_particle_create(int(particles), int(species), stream, C.byref(self._h)), "mdsim.particle")
        self.nparticle = int(particles)
        self.nspecies = int(species)
        self.array_size = lib.hxm_particle_array_size(self._h)

    def _shape(self, name):
        if name not in self._arrays:
            raise HxError(f"particle: unknown array '{name}'")
        dtype, ncomp = self._arrays[name]
        return dtype, (self.nparticle, ncomp) if ncomp else (self.nparticle,)

    def get(self, name, out=None):
        """Host copy of a particle array; `out` (C-contiguous, right dtype and shape — e.g. a view of pinned memory, which
        lets the device-to-host copy run at the full PCIe rate) is filled and returned when given."""
        dtype, shape = self._shape(name)
        if out is None:
            out = np.zeros(shape, dtype=dtype)
        elif out.dtype != dtype or out.shape != shape or not out.flags["C_CONTIGUOUS"]:
            raise HxError(f"particle.get('{name}'): out must be a C-contiguous {np.dtype(dtype).name} array of shape {shape}")
        check(lib.hxm_particle_get(self._h, name.encode(), out.ctypes.data_as(C.c_void_p)), f"particle.get('{name}')")
        return out

    def set(self, name, value):
        dtype, shape = self._shape(name)
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(value, dtype=dtype), shape))
        check(lib.hxm_particle_set(self._h, name.encode(), a.ctypes.data_as(C.c_void_p)), f"particle.set('{name}')")

    # Lua: particle.data[name]
    @property
    def data(self):
        p = self

        class _Data:
            def __getitem__(self, name):
                return p.get(name)

            def __setitem__(self, name, value):
                p.set(name, value)
        return _Data()

    def device_ptr(self, name):
        return lib.hxm_particle_device_ptr(self._h, name.encode())

    def aux_enable(self):
        check(lib.hxm_particle_aux_enable(self._h))


# ---- potentials (lua/halmd/mdsim/potentials/pair/lennard_jones.lua.in, truncate.lua.in) ---------------------------
class _lennard_jones:
    def __init__(self, epsilon=1.0, sigma=1.0, species=None):
        eps = np.asarray(epsilon, dtype=np.float32)
        n = species if species is not None else (eps.shape[0] if eps.ndim == 2 else 1)
        self.species = int(n)
        self.epsilon = _matrix(epsilon, n, n, "epsilon")
        self.sigma = _matrix(sigma, n, n, "sigma")
        self.r_core_sigma = None

    def modify(self, *args, **kw):
        """potential:modify({"hard_core", radius = ...}) (adapters/modify.lua.in): the potential is evaluated at
        r - radius * sigma.  Returns a new potential that can be truncated."""
        spec = dict(args[0]) if args and isinstance(args[0], dict) else dict(kw)
        kind = spec.pop("type", None) or (args[0] if args and isinstance(args[0], str) else None)
        if kind != "hard_core":
            raise ValueError(f"invalid modification type '{kind}'")
        if "radius" not in spec and "radius" not in kw:
            raise ValueError("missing argument 'radius'")
        radius = spec.get("radius", kw.get("radius"))
        new = _lennard_jones(self.epsilon, self.sigma, self.species)
        new.r_core_sigma = _matrix(radius, self.species, self.species, "radius")
        return new

    def truncate(self, type=None, cutoff=None, h=None, **kw):
        """potential:truncate{[type], cutoff=..., h=...}; the default truncation is 'shifted' (truncate.lua.in:289)."""
        if isinstance(type, dict):
            kw, type = type, type.get("type")
        cutoff = kw.get("cutoff", cutoff)
        h = kw.get("h", h)
        kind = type or kw.get("type") or "shifted"
        if cutoff is None:
            raise ValueError("truncate: missing argument 'cutoff'")
        kinds = {"sharp": hx.TRUNC_SHARP, "shifted": hx.TRUNC_SHIFTED, "smooth_r4": hx.TRUNC_SMOOTH_R4,
                 "force_shifted": hx.TRUNC_FORCE_SHIFTED}
        if kind not in kinds:
            raise ValueError(f"truncate: truncation '{kind}' is not built (available: {sorted(kinds)})")
        if kind == "smooth_r4" and h is None:
            raise ValueError("truncate: smooth_r4 requires 'h'")
        return _truncated(self, kinds[kind], kind, _matrix(cutoff, self.species, self.species, "cutoff"), float(h or 0.0))


class _truncated(_Handle):
    _destroy = "hxm_potential_destroy"

    def __init__(self, parent, trunc, name, cutoff, h):
        super().__init__()
        self.parent, self.truncation, self.cutoff, self.h = parent, name, cutoff, h
        self.species = parent.species
        self.epsilon, self.sigma = parent.epsilon, parent.sigma
        n = parent.species
        self.r_core_sigma = parent.r_core_sigma
        if parent.r_core_sigma is not None:
            check(lib.hxm_potential_lj_hard_core_create(n, n, hx.farr(parent.epsilon.ravel()), hx.farr(parent.sigma.ravel()),
                                                        hx.farr(parent.r_core_sigma.ravel()), trunc,
                                                        hx.farr(cutoff.ravel()), h, C.byref(self._h)), "potential")
        else:
            check(lib.hxm_potential_lj_create(n, n, hx.farr(parent.epsilon.ravel()), hx.farr(parent.sigma.ravel()), trunc,
                                              hx.farr(cutoff.ravel()), h, C.byref(self._h)), "potential")
        out = (C.c_float * (n * n))()
        check(lib.hxm_potential_r_cut(self._h, out))
        self.r_cut = np.array(list(out), dtype=np.float32).reshape(n, n)
        self.r_cut_sigma = cutoff

    @property
    def params(self):
        return lib.hxm_potential_params(self._h).contents


# ---- mdsim.binning ------------------------------------------------------------------------------------------------
class binning(_Handle):
    _destroy = "hxm_binning_destroy"

    def __init__(self, box=None, particle=None, r_cut=None, skin=0.5, occupancy=None):
        super().__init__()
        n = particle.nspecies
        rc = _matrix(r_cut, n, n, "r_cut")
        self.box, self.particle, self.r_skin = box, particle, skin
        check(lib.hxm_binning_create(particle._h, box._d3(), hx.farr(rc.ravel()), n, n, float(skin),
                                     float(occupancy or 0.0), C.byref(self._h)), "mdsim.binning")

    @property
    def ncell(self):
        out = (C.c_uint * 3)()
        check(lib.hxm_binning_ncell(self._h, out))
        return list(out)

    @property
    def cell_size(self):
        return lib.hxm_binning_cell_size(self._h)

    def g_cell(self):
        """binning:g_cell() — (ncells, cell_size) uint32, placeholders 0xFFFFFFFF."""
        check(lib.hxm_binning_update(self._h, None))
        ncells = int(np.prod(self.ncell))
        out = np.zeros((ncells, self.cell_size), dtype=np.uint32)
        check(lib.hxm_binning_update(self._h, out.ctypes.data_as(C.c_void_p)))
        return out


# ---- mdsim.neighbour (lua/halmd/mdsim/neighbour.lua.in) -----------------------------------------------------------------
class neighbour(_Handle):
    _destroy = "hxm_neighbour_destroy"

    def __init__(self, particle=None, box=None, r_cut=None, skin=0.5, occupancy=None, algorithm=None,
                 unroll_force_loop=None, disable_binning=False, disable_sorting=False, binning=None, layout=None):
        super().__init__()
        if disable_binning:
            raise HxError("neighbour: disable_binning (neighbours.from_particle) is out of scope of this build")
        if isinstance(particle, (list, tuple)):
            if particle[0] is not particle[1]:
                raise HxError("neighbour: distinct particle instances are not built")
            particle = particle[0]
        n = particle.nspecies
        rc = _matrix(r_cut, n, n, "r_cut")
        # The reference offers a column-major list (default) and a row-major one (unroll_force_loop); here the
        # row-major layout with 128-bit coalesced reads is the default and `unroll_force_loop=False` plus
        # layout="column" selects the reference's default layout.
        if layout is None:
            layout = "row"
        self.layout = {"column": hx.LAYOUT_COLUMN, "row": hx.LAYOUT_ROW}[layout]
        self.particle, self.box, self.r_skin, self.binning = particle, box, skin, binning
        check(lib.hxm_neighbour_create(particle._h, box._d3(), hx.farr(rc.ravel()), n, n, float(skin),
                                       float(occupancy or 0.0), self.layout, int(bool(disable_sorting)),
                                       binning._h if binning is not None else None, C.byref(self._h)), "mdsim.neighbour")

    size = property(lambda self: lib.hxm_neighbour_size(self._h))
    stride = property(lambda self: lib.hxm_neighbour_stride(self._h))
    updates = property(lambda self: lib.hxm_neighbour_updates(self._h))

    def check(self):
        r = C.c_int()
        check(lib.hxm_neighbour_check(self._h, C.byref(r)), "neighbour.check")
        return bool(r.value)

    def mean_length(self):
        return lib.hxm_neighbour_mean_length(self._h)

    def lists(self):
        """Per-particle neighbour index arrays (memory order), after updating the lists if necessary."""
        self.check()
        n = self.particle.nparticle
        words = n * self.stride if self.layout == hx.LAYOUT_ROW else self.size * self.stride
        raw = np.zeros(words, dtype=np.uint32)
        cnt = np.zeros(n, dtype=np.uint32)
        check(lib.hxm_neighbour_get(self._h, raw.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.c_void_p)))
        if self.layout == hx.LAYOUT_ROW:
            rows = raw.reshape(n, self.stride)
            return [rows[i, :cnt[i]] for i in range(n)], cnt
        cols = raw.reshape(self.size, self.stride)
        return [cols[:cnt[i], i] for i in range(n)], cnt


# ---- mdsim.forces.pair (lua/halmd/mdsim/forces/pair.lua.in → pair_trunc.lua.in) ----------------------------------------
class _pair_force(_Handle):
    _destroy = "hxm_force_destroy"

    def __init__(self, box=None, particle=None, potential=None, neighbour=None, weight=1.0, **neighbour_args):
        super().__init__()
        if isinstance(particle, (list, tuple)):
            particle = particle[0]
        if not isinstance(potential, _truncated):
            raise HxError("forces.pair: untruncated potentials (pair_full) are out of scope; use potential:truncate{...}")
        if neighbour is None:
            # pair_trunc.lua.in:141 — construct the neighbour module from the potential's cutoff
            neighbour = globals()["neighbour"](particle=particle, box=box, r_cut=potential.r_cut, **neighbour_args)
        self.box, self.particle, self.potential, self.neighbour = box, particle, potential, neighbour
        check(lib.hxm_force_pair_trunc_create(potential._h, particle._h, box._d3(), neighbour._h, float(weight),
                                              C.byref(self._h)), "mdsim.forces.pair")

    def apply(self):
        check(lib.hxm_force_apply(self._h), "force.apply")


# ---- mdsim.potentials.external.planar_wall / mdsim.forces.external ---------------------------------------------------------
class _planar_wall:
    """lua/halmd/mdsim/potentials/external/planar_wall.lua.in: offset[nwall], surface_normal[nwall][3] (or one normal
    for all walls), epsilon / sigma / wetting / cutoff as numbers or [nwall][species] matrices, smoothing, species."""

    def __init__(self, offset=None, surface_normal=None, epsilon=None, sigma=None, wetting=None, cutoff=None,
                 smoothing=None, species=None, memory=None, label=None):
        for name, v in (("offset", offset), ("surface_normal", surface_normal), ("epsilon", epsilon), ("sigma", sigma),
                        ("wetting", wetting), ("cutoff", cutoff), ("smoothing", smoothing), ("species", species)):
            if v is None:
                raise ValueError(f"missing argument '{name}'")
        offset = np.atleast_1d(np.asarray(offset, dtype=np.float32))
        nwall, species = len(offset), int(species)
        normal = np.asarray(surface_normal, dtype=np.float32)
        if normal.ndim == 1:
            normal = np.tile(normal, (nwall, 1))
        if normal.shape != (nwall, 3):
            raise ValueError("bad argument 'surface_normal'")

        def matrix(v, name):
            m = np.asarray(v, dtype=np.float32)
            if m.ndim == 0:
                m = np.full((nwall, species), float(v), dtype=np.float32)
            if m.shape != (nwall, species):
                raise ValueError(f"mismatching shape of argument '{name}'")
            return np.ascontiguousarray(m)

        self.offset, self.surface_normal = offset, normal
        self.epsilon, self.sigma = matrix(epsilon, "epsilon"), matrix(sigma, "sigma")
        self.wetting, self.cutoff = matrix(wetting, "wetting"), matrix(cutoff, "cutoff")
        self.smoothing, self.species, self.nwall = float(smoothing), species, nwall
        self.params = hx.PlanarWall()
        check(lib.hx_planar_wall_init(C.byref(self.params), nwall, species, hx.farr(offset), hx.farr(normal.reshape(-1)),
                                      hx.farr(self.epsilon.reshape(-1)), hx.farr(self.sigma.reshape(-1)),
                                      hx.farr(self.wetting.reshape(-1)), hx.farr(self.cutoff.reshape(-1)), self.smoothing),
              "potentials.external.planar_wall")


class _external_force:
    """lua/halmd/mdsim/forces/external.lua.in.  The contribution is evaluated after the particle's pair force (which
    must have been constructed before), adding to the same force array."""

    def __init__(self, box=None, particle=None, potential=None):
        if not isinstance(potential, _planar_wall):
            raise HxError("forces.external: only potentials.external.planar_wall is supported")
        pair = getattr(particle, "_force", None)
        if pair is None:
            raise HxError("forces.external: construct mdsim.forces.pair{particle=...} first")
        self.box, self.particle, self.potential, self.pair = box, particle, potential, pair
        check(lib.hxm_force_add_planar_wall(pair._h, C.byref(potential.params)), "mdsim.forces.external")

    def apply(self):
        self.pair.apply()


# ---- integrators -----------------------------------------------------------------------------------------------------------
class _integrator(_Handle):
    _destroy = "hxm_integrator_destroy"

    def integrate(self):
        check(lib.hxm_integrator_integrate(self._h), "integrator.integrate")

    def finalize(self):
        check(lib.hxm_integrator_finalize(self._h), "integrator.finalize")

    def set_timestep(self, timestep):
        check(lib.hxm_integrator_set_timestep(self._h, float(timestep)))
        self.timestep = float(timestep)


def _find_force(particle):
    f = getattr(particle, "_force", None)
    if f is None:
        raise HxError("integrator: construct mdsim.forces.pair{particle=...} before the integrator")
    return f


class _verlet(_integrator):
    def __init__(self, box=None, particle=None, timestep=0.001, force=None):
        super().__init__()
        force = force or _find_force(particle)
        self.box, self.particle, self.force, self.timestep = box, particle, force, float(timestep)
        check(lib.hxm_integrator_verlet_create(particle._h, force._h, box._d3(), float(timestep), C.byref(self._h)),
              "integrators.verlet")
        core._register(self)


class _verlet_nvt_andersen(_integrator):
    def __init__(self, box=None, particle=None, timestep=0.001, temperature=1.0, rate=None, seed=0, force=None):
        super().__init__()
        force = force or _find_force(particle)
        if rate is None:
            raise ValueError("missing argument 'rate'")
        self.box, self.particle, self.force, self.timestep = box, particle, force, float(timestep)
        self.temperature, self.rate = float(temperature), float(rate)
        check(lib.hxm_integrator_andersen_create(particle._h, force._h, box._d3(), float(timestep), float(temperature),
                                                 float(rate), int(seed), C.byref(self._h)), "integrators.verlet_nvt_andersen")
        core._register(self)

    def set_temperature(self, temperature):
        check(lib.hxm_integrator_set_temperature(self._h, float(temperature)))
        self.temperature = float(temperature)


class _verlet_nvt_hoover(_integrator):
    """mdsim.integrators.verlet_nvt_hoover (lua/halmd/mdsim/integrators/verlet_nvt_hoover.lua.in:57-61): Nosé–Hoover
    chain of length 2; `position`, `velocity`, `mass` are the chain variables, `internal_energy` the chain's energy
    per particle, so that E_tot + internal_energy is conserved."""

    def __init__(self, box=None, particle=None, timestep=0.001, temperature=1.0, resonance_frequency=5.0, force=None):
        super().__init__()
        force = force or _find_force(particle)
        self.box, self.particle, self.force, self.timestep = box, particle, force, float(timestep)
        self.temperature, self.resonance_frequency = float(temperature), float(resonance_frequency)
        check(lib.hxm_integrator_hoover_create(particle._h, force._h, box._d3(), float(timestep), float(temperature),
                                               float(resonance_frequency), C.byref(self._h)),
              "integrators.verlet_nvt_hoover")
        core._register(self)

    def _state(self):
        out = (C.c_double * 7)()
        check(lib.hxm_integrator_hoover_state(self._h, out), "integrators.verlet_nvt_hoover")
        return list(out)

    position = property(lambda self: self._state()[0:2])
    velocity = property(lambda self: self._state()[2:4])
    mass = property(lambda self: self._state()[4:6])
    internal_energy = property(lambda self: self._state()[6])

    def set_temperature(self, temperature):
        check(lib.hxm_integrator_set_temperature(self._h, float(temperature)))
        self.temperature = float(temperature)

    def set_mass(self, mass):
        check(lib.hxm_integrator_hoover_set_mass(self._h, (C.c_double * 2)(*map(float, mass))))


# ---- mdsim.core + observables.sampler ------------------------------------------------------------------------------------------
class _core(_Handle):
    _destroy = "hxm_core_destroy"

    def __init__(self, integrator):
        super().__init__()
        self.integrator = integrator
        check(lib.hxm_core_create(integrator._h, C.byref(self._h)), "mdsim.core")

    def mdstep(self):
        check(lib.hxm_core_mdstep(self._h), "core.mdstep")

    def run(self, steps):
        check(lib.hxm_core_run(self._h, int(steps)), "sampler.run")

    @property
    def step(self):
        return lib.hxm_core_step(self._h)

    def set_chunk(self, lo, hi):
        check(lib.hxm_core_set_chunk(self._h, int(lo), int(hi)))

    def set_fused_verlet(self, enable):
        check(lib.hxm_core_set_fused_verlet(self._h, 1 if enable else 0))

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib.hxm_core_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"launches": a.value, "chunks": b.value, "syncs": c.value}

    def thermodynamics(self):
        out = (C.c_double * 10)()
        check(lib.hxm_core_thermodynamics(self._h, out), "thermodynamics")
        keys = ["kinetic_energy", "potential_energy", "virial", "temperature", "pressure", "internal_energy"]
        d = dict(zip(keys, list(out)[:6]))
        d["center_of_mass_velocity"] = list(out)[6:9]
        d["density"] = out[9]
        return d


class core:
    """mdsim.core singleton semantics: the most recently constructed integrator drives the step loop."""
    _current = None

    @classmethod
    def _register(cls, integrator):
        cls._current = _core(integrator)

    @classmethod
    def get(cls):
        if cls._current is None:
            raise HxError("mdsim.core: no integrator has been constructed")
        return cls._current


class _sampler:
    """observables.sampler — sampler:run(steps) (halmd/observables/sampler.cpp:49-83)."""

    def run(self, steps, blocking_scheme=None):
        c = core.get()
        if blocking_scheme is None:
            c.run(steps)
        else:
            check(lib.hxm_core_run_sampling(c._h, int(steps), blocking_scheme._h), "sampler.run")

    @property
    def step(self):
        return core.get().step


class _thermodynamics:
    """observables.thermodynamics{box, group, force} — scalar observables of the whole system."""

    def __init__(self, box=None, group=None, force=None, particle=None):
        self.box = box

    def __getattr__(self, name):
        d = core.get().thermodynamics()
        if name in d:
            return lambda: d[name]
        if name == "total_energy":
            return lambda: d["internal_energy"]
        raise AttributeError(name)

    def all(self):
        return core.get().thermodynamics()


class _semilog_grid:
    """observables.utility.semilog_grid{start, stop, decimation} (halmd/observables/utility/semilog_grid.cpp:31-52):
    linear grid from `start` with spacing `start`, the spacing doubled after every `decimation` points (0: never)."""

    def __init__(self, start=None, stop=None, decimation=0):
        grid, h, x, i = [], float(start), float(start), 0
        while x < stop:
            grid.append(x)
            x += h
            i += 1
            if decimation > 0 and i % decimation == 0:
                h *= 2
        self.value = grid


class _wavevector:
    """observables.utility.wavevector{box, wavenumber, tolerance, max_count} — reciprocal lattice vectors grouped in
    shells (halmd/observables/utility/wavevector.cpp:40-120, halmd/algorithm/host/pick_lattice_points.hpp:62-160)."""

    def __init__(self, box=None, wavenumber=None, tolerance=0.01, max_count=7, filter=None):
        wavenumber = [float(q) for q in wavenumber]
        nw = len(wavenumber)
        cap = nw * int(max_count)
        vec = np.zeros((max(cap, 1), 3))
        offs = (C.c_uint * (nw + 1))()
        filt = None if filter is None else (C.c_uint * 3)(*[int(x) for x in filter])
        check(lib.hx_wavevector_pick((C.c_double * nw)(*wavenumber), nw, box._d3(), float(tolerance), int(max_count), filt,
                                     vec.ctypes.data_as(hx.D3), cap, offs), "observables.utility.wavevector")
        offs = list(offs)
        keep = [i for i in range(nw) if offs[i + 1] > offs[i]]       # wavevector.cpp:103-118: empty shells are discarded
        self.wavenumber = [wavenumber[i] for i in keep]
        self.shell, rows = [], []
        for i in keep:
            self.shell.append((len(rows), len(rows) + offs[i + 1] - offs[i]))
            rows.extend(range(offs[i], offs[i + 1]))
        self.value = np.ascontiguousarray(vec[rows])


class _density_mode:
    """observables.density_mode{group / particle, wavevector} (lua/halmd/observables/density_mode.lua.in):
    rho_q = sum_j exp(-i q.r_j) for every wavevector, as a complex array."""

    def __init__(self, group=None, particle=None, wavevector=None):
        self.particle = particle if particle is not None else getattr(group, "particle", group)
        if self.particle is None or wavevector is None:
            raise ValueError("density_mode: missing argument 'group' or 'wavevector'")
        self.wavevector = wavevector

    def acquire(self):
        q = np.ascontiguousarray(self.wavevector.value, dtype=np.float32)
        rho = np.zeros(2 * len(q))
        check(lib.hxm_density_modes(self.particle._h, hx.farr(q.reshape(-1)), len(q), rho.ctypes.data_as(C.c_void_p)),
              "density_mode.acquire")
        return rho[0::2] + 1j * rho[1::2]


class _ssf:
    """observables.ssf{density_mode, norm} (lua/halmd/observables/ssf.lua.in, halmd/observables/ssf.cpp:58-105): per
    wavenumber shell, mean / error of mean / count of Re[rho1_q conj(rho2_q)] / norm.  `density_mode` is one module
    or a pair (partial structure factors)."""

    def __init__(self, density_mode=None, norm=None):
        if density_mode is None or norm is None:
            raise ValueError("ssf: missing argument 'density_mode' or 'norm'")
        modes = density_mode if isinstance(density_mode, (list, tuple)) else (density_mode, density_mode)
        self.mode1, self.mode2 = modes[0], modes[-1]
        self.wavevector = self.mode1.wavevector
        self.norm = float(norm)

    def sample(self):
        rho1 = self.mode1.acquire()
        rho2 = rho1 if self.mode2 is self.mode1 else self.mode2.acquire()
        out = np.zeros((len(self.wavevector.shell), 3))
        for k, (a, b) in enumerate(self.wavevector.shell):
            re = (rho1[a:b] * np.conj(rho2[a:b])).real / self.norm
            n = b - a
            # accumulator.hpp:139-180 — error of mean = sqrt(variance / (count - 1)), variance = m2 / count
            err = np.sqrt(re.var() / (n - 1)) if n > 1 else 0.0
            out[k] = (re.mean(), err, n)
        return out


class _blocking_scheme(_Handle):
    """observables.dynamics.blocking_scheme{max_lag, every, size, shift, separation} with the correlation functions
    mean_square_displacement and (given a wavevector module) intermediate_scattering_function."""
    _destroy = "hxm_blocking_destroy"

    def __init__(self, max_lag=None, every=None, size=10, shift=None, separation=None, wavevector=None):
        super().__init__()
        c = core.get()
        # the C object refers to the core and, through it, to the integrator: keep both alive as long as this handle
        self._core, self._integrator = c, c.integrator
        dt = c.integrator.timestep
        shift = int(np.floor(np.sqrt(size))) if shift is None else int(shift)        # blocking_scheme.lua.in
        separation = size if separation is None else int(separation)
        self.size, self.wavevector = int(size), wavevector
        if wavevector is not None:
            q = np.ascontiguousarray(wavevector.value, dtype=np.float32)
            offs = [s[0] for s in wavevector.shell] + [wavevector.shell[-1][1]]
            check(lib.hxm_blocking_create(c._h, float(max_lag), float(every) * dt, int(size), shift, separation,
                                          q.ctypes.data_as(hx.F3), len(q), hx.u3(offs) if len(offs) == 3 else (C.c_uint * len(offs))(*offs),
                                          len(wavevector.shell), C.byref(self._h)), "blocking_scheme")
        else:
            check(lib.hxm_blocking_create(c._h, float(max_lag), float(every) * dt, int(size), shift, separation, None, 0,
                                          None, 0, C.byref(self._h)), "blocking_scheme")
        self.levels = lib.hxm_blocking_levels(self._h)
        iv = (C.c_uint64 * self.levels)()
        check(lib.hxm_blocking_intervals(self._h, iv))
        self.interval = list(iv)
        self.time = np.array([[i * j * dt for j in range(self.size)] for i in self.interval])

    def sample(self):
        check(lib.hxm_blocking_sample(self._h))

    def finalise(self):
        check(lib.hxm_blocking_finalise(self._h))

    def trace(self, enable=True):
        """Parity diagnostics: start (or stop and clear) recording, and return the (level, first step, later step) triples
        of all sample pairs correlated so far, in the order correlation<>::compute visits them."""
        n = C.c_size_t(0)
        check(lib.hxm_blocking_trace(self._h, 1, None, 0, C.byref(n)))
        out = np.zeros((n.value, 3), dtype=np.uint64)
        check(lib.hxm_blocking_trace(self._h, 1 if enable else 0, out.ctypes.data_as(C.POINTER(C.c_uint64)), n.value, None))
        return out

    def _result(self, fn, extra):
        shape = (self.levels, self.size) + extra + (3,)
        out = np.zeros(shape)
        check(fn(self._h, out.ctypes.data_as(hx.D3)))
        cnt, mean, m2 = out[..., 0], out[..., 1], out[..., 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            err = np.sqrt(m2 / cnt / np.maximum(cnt - 1, 1))       # error_of_mean (accumulator.hpp:170-180)
        return {"count": cnt, "mean": mean, "error": err}

    def mean_square_displacement(self):
        return self._result(lib.hxm_blocking_msd, ())

    def mean_quartic_displacement(self):
        return self._result(lib.hxm_blocking_mqd, ())

    def velocity_autocorrelation(self):
        return self._result(lib.hxm_blocking_vacf, ())

    def intermediate_scattering_function(self):
        if self.wavevector is None:
            raise HxError("blocking_scheme: no wavevector module given")
        return self._result(lib.hxm_blocking_isf, (len(self.wavevector.shell),))


# ---- namespaces as in Lua ----------------------------------------------------------------------------------------------------------
class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _pair(**kw):
    f = _pair_force(**kw)
    f.particle._force = f
    return f


mdsim = _NS(
    box=box, particle=particle, binning=binning, neighbour=neighbour, core=core,
    potentials=_NS(pair=_NS(lennard_jones=_lennard_jones), external=_NS(planar_wall=_planar_wall)),
    forces=_NS(pair=_pair, external=_external_force),
    integrators=_NS(verlet=_verlet, verlet_nvt_andersen=_verlet_nvt_andersen, verlet_nvt_hoover=_verlet_nvt_hoover),
)

observables = _NS(
    sampler=_sampler(),
    thermodynamics=_thermodynamics,
    utility=_NS(wavevector=_wavevector, semilog_grid=_semilog_grid),
    density_mode=_density_mode,
    ssf=_ssf,
    dynamics=_NS(blocking_scheme=_blocking_scheme),
)
