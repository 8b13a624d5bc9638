"""Parameter builders of the reference, same names and semantics.

StokesParams / StokesParams1DBuilder / StokesParams1D ............ solvers/fem/stokes_solver/mod.rs:14,27-31,55-119, dim1.rs:26-52
DiffussionParams / ...TimeIndependentBuilder / ...TimeDependentBuilder . solvers/fem/diffusion_solver/mod.rs:11-181
Builders are immutable-style (each setter returns a new builder, like the Rust `Self { x: Some(x), ..self }`), and
`build()` raises (the reference `panic!`s) when a mandatory field is missing, with the reference's messages."""
from dataclasses import dataclass, field, replace
from typing import Callable, List, Optional, Tuple


class BuilderPanic(RuntimeError):
    """Stands in for the `panic!` of the reference's `build()` methods."""


# ---------------------------------------------------------------- Stokes
@dataclass(frozen=True)
class StokesParams1D:                      # dim1.rs:26-30
    rho: float
    hydrostatic_pressure: float
    force_function: Callable[[float], float]

    def __repr__(self):                    # dim1.rs:43-50 (Debug prints f(0))
        return ("{ rho: %r,\nhydrostatic_pressure: %r,\n force_function: f(0) -> %r }"
                % (self.rho, self.hydrostatic_pressure, self.force_function(0.0)))


StaticPressureParams = StokesParams1D      # stokes_solver/mod.rs:10


@dataclass(frozen=True)
class StokesParams1DBuilder:               # stokes_solver/mod.rs:27-31
    _hydrostatic_pressure: Optional[float] = None
    _rho: Optional[float] = None
    _force_function: Optional[Callable[[float], float]] = None

    def hydrostatic_pressure(self, pressure_value: float) -> "StokesParams1DBuilder":   # :72-77
        return replace(self, _hydrostatic_pressure=float(pressure_value))

    def density(self, rho: float) -> "StokesParams1DBuilder":                           # :79-84
        return replace(self, _rho=float(rho))

    def force_function(self, func: Callable[[float], float]) -> "StokesParams1DBuilder":  # :86-91
        return replace(self, _force_function=func)

    def build(self) -> StokesParams1D:                                                  # :93-118
        if self._hydrostatic_pressure is None:
            raise BuilderPanic("Params lack 'pressure' term!")
        if self._rho is None:
            raise BuilderPanic("Params lack 'density' term!")
        if self._force_function is None:
            raise BuilderPanic("Params lack force_function!")
        return StokesParams1D(self._rho, self._hydrostatic_pressure, self._force_function)


StaticPressureParamsBuilder = StokesParams1DBuilder  # stokes_solver/mod.rs:11


class StokesParams:                        # stokes_solver/mod.rs:14,55-68
    @staticmethod
    def normal_1d() -> StokesParams1DBuilder:
        return StokesParams1DBuilder()

    @staticmethod
    def static_pressure() -> StokesParams1DBuilder:
        return StokesParams1DBuilder()

    @staticmethod
    def normal_2d():
        # the reference's 2D solver is an empty stub whose run() panics "Not implemented yet!" (dzahui_window.rs:792-794)
        raise NotImplementedError("Not implemented yet!")


# ---------------------------------------------------------------- diffusion
@dataclass(frozen=True)
class DiffussionParamsTimeIndependent:     # time_independent.rs:26-30
    mu: float
    b: float
    boundary_conditions: Tuple[float, float]


@dataclass(frozen=True)
class DiffussionParamsTimeDependent:       # time_dependent.rs:25-30
    mu: float
    b: float
    boundary_conditions: Tuple[float, float]
    initial_conditions: List[float] = field(default_factory=list)


@dataclass(frozen=True)
class DiffussionParamsTimeIndependentBuilder:   # diffusion_solver/mod.rs:43-47
    _mu: Optional[float] = None
    _b: Optional[float] = None
    _boundary_conditions: Optional[Tuple[float, float]] = None

    def mu(self, mu: float):
        return replace(self, _mu=float(mu))

    def b(self, b: float):
        return replace(self, _b=float(b))

    def boundary_conditions(self, left: float, right: float):
        return replace(self, _boundary_conditions=(float(left), float(right)))

    def build(self) -> DiffussionParamsTimeIndependent:   # :154-180
        if self._mu is None:
            raise BuilderPanic("Params lack 'mu' term!")
        if self._b is None:
            raise BuilderPanic("Params lack 'b' term!")
        if self._boundary_conditions is None:
            raise BuilderPanic("Params lack boundary conditions!")
        return DiffussionParamsTimeIndependent(self._mu, self._b, self._boundary_conditions)


@dataclass(frozen=True)
class DiffussionParamsTimeDependentBuilder:     # diffusion_solver/mod.rs:25-30
    _mu: Optional[float] = None
    _b: Optional[float] = None
    _boundary_conditions: Optional[Tuple[float, float]] = None
    _initial_conditions: Optional[List[float]] = None

    def b(self, b: float):
        return replace(self, _b=float(b))

    def mu(self, mu: float):
        return replace(self, _mu=float(mu))

    def boundary_conditions(self, left: float, right: float):
        return replace(self, _boundary_conditions=(float(left), float(right)))

    def initial_conditions(self, initial_conditions):   # :82-87
        return replace(self, _initial_conditions=[float(v) for v in initial_conditions])

    @staticmethod
    def initial_conditions_from_function(_func, _mesh):  # :89-95 is `todo!()` in the reference
        raise NotImplementedError("not yet implemented")

    def build(self) -> DiffussionParamsTimeDependent:   # :97-129
        if self._mu is None:
            raise BuilderPanic("Params lack 'mu' term!")
        if self._b is None:
            raise BuilderPanic("Params lack 'b' term!")
        if self._boundary_conditions is None:
            raise BuilderPanic("Params lack boundary conditions!")
        if self._initial_conditions is None:
            raise BuilderPanic("Params lack initial conditions!")
        return DiffussionParamsTimeDependent(self._mu, self._b, self._boundary_conditions, list(self._initial_conditions))


class DiffussionParams:                    # diffusion_solver/mod.rs:11,50-58
    @staticmethod
    def time_dependent() -> DiffussionParamsTimeDependentBuilder:
        return DiffussionParamsTimeDependentBuilder()

    @staticmethod
    def time_independent() -> DiffussionParamsTimeIndependentBuilder:
        return DiffussionParamsTimeIndependentBuilder()
