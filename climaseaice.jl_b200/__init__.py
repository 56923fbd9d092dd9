"""climaseaice_b200 -- B200-native drop-in for ClimaSeaIce.jl's EVP-substep + h/aice advection hot path.

The product is csrc/ -> libclimaseaice_b200.so (C ABI in include/climaseaice_b200.h).  The Python
modules here are the host-side mirror of the reference's user interface for that path.
"""
from . import _lib
from ._lib import CsiError, lib
from .model import (Bounded, Folded, Center, ConductiveFlux, ElastoViscoPlasticRheology, FPlane, Face, Field, Flat, HydrostaticSphericalCoriolis,
                    IceWaterThermalEquilibrium, LatitudeLongitudeGrid, LinearHeatFlux, OrthogonalSphericalShellGrid, MeltingConstrainedFluxBalance, Periodic,
                    PhaseTransitions, PrescribedTemperature, RadiativeEmission, RectilinearGrid, SeaIceModel,
                    SeaIceMomentumEquation, SemiImplicitStress, SlabThermodynamics, SplitExplicitSolver,
                    StressBalanceFreeDrift, UpwindBiased, ValueBoundaryCondition, WENO, nccl_unique_id,
                    sea_ice_slab_thermodynamics, snow_slab_thermodynamics, time_step_b)

__all__ = [n for n in dir() if not n.startswith("_")]
