"""Horizontal metrics of a regularly spaced LatitudeLongitudeGrid (numpy only; shared by the host mirror and the
synthetic cases).  They depend on j only: dx = R cos(phi) dlambda at centre / face latitudes, dy = R dphi,
Az = R^2 dlambda (sin phi_north - sin phi_south) -- Oceananigans' precomputed lat-lon metrics."""
from __future__ import annotations

import numpy as np

METRIC_NAMES = ("dxcc", "dxfc", "dxcf", "dxff", "dycc", "dyfc", "dycf", "dyff", "azcc", "azfc", "azcf", "azff")


def latitude_longitude_metrics(Nx, Ny, Hy, longitude, latitude, radius=6371e3):
    """dict name -> array of Ny + 2Hy + 1 doubles, entry for index j (1-based, halos included) at [j - 1 + Hy]."""
    R = float(radius)
    dl = np.deg2rad((longitude[1] - longitude[0]) / Nx)
    dp = (latitude[1] - latitude[0]) / Ny
    j = np.arange(1 - Hy, Ny + Hy + 2)
    p0 = latitude[0]
    phif, phic = np.deg2rad(p0 + (j - 1) * dp), np.deg2rad(p0 + (j - 0.5) * dp)
    phif_n, phic_s = np.deg2rad(p0 + j * dp), np.deg2rad(p0 + (j - 1.5) * dp)
    dxc, dxf = R * np.cos(phic) * dl, R * np.cos(phif) * dl
    dy = np.full_like(dxc, R * np.deg2rad(dp))
    azc, azf = R * R * dl * (np.sin(phif_n) - np.sin(phif)), R * R * dl * (np.sin(phic) - np.sin(phic_s))
    return dict(dxcc=dxc, dxfc=dxc, dxcf=dxf, dxff=dxf, dycc=dy, dyfc=dy, dycf=dy, dyff=dy, azcc=azc, azfc=azc, azcf=azf, azff=azf)


def spherical_coriolis_f_ff(Ny, Hy, latitude, rotation_rate=7.292115e-5):
    """f at (Face, Face) = 2 Omega sin(phi_f) per row (HydrostaticSphericalCoriolis), row j at [j - 1 + Hy]."""
    dp = (latitude[1] - latitude[0]) / Ny
    j = np.arange(1 - Hy, Ny + Hy + 2)
    return np.ascontiguousarray(2 * rotation_rate * np.sin(np.deg2rad(latitude[0] + (j - 1) * dp)))
