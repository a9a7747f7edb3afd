"""Shared fixtures for the parity tests: small synthetic MPAS-format cases (seeded, deterministic)."""
from __future__ import annotations

import functools

import numpy as np

from mops_b200 import synthetic as S


@functools.lru_cache(maxsize=None)
def mesh(level: int):
    return S.icosahedral_mesh(level)


@functools.lru_cache(maxsize=None)
def snapshots(level: int, n_levels: int, variant: str):
    """variant 'plain': BASELINE-style (uniform layers, same velocity in all layers, w = 0);
    'rich': sheared + bumpy bathymetry + vertical velocity + two scalar attributes."""
    m = mesh(level)
    if variant == "plain":
        return (S.solid_body_snapshot(m, n_levels, 0.5, tilt=0.3),
                S.solid_body_snapshot(m, n_levels, 0.6, tilt=0.31))
    if variant == "rich":
        return (S.solid_body_snapshot(m, n_levels, 2.0, tilt=0.3, shear=0.4, bumpy=0.3, w_amp=2e-3, with_attrs=True),
                S.solid_body_snapshot(m, n_levels, 3.0, tilt=0.35, shear=0.2, bumpy=0.25, w_amp=-1e-3, with_attrs=True))
    if variant == "nonmono":
        # a few cells get a negative layer thickness -> non-monotone columns -> full-column path
        a, b = snapshots(level, n_levels, "rich")
        import copy
        a, b = copy.deepcopy(a), copy.deepcopy(b)
        rng = np.random.default_rng(7)
        for s in (a, b):
            idx = rng.choice(m.n_cells, size=max(4, m.n_cells // 10), replace=False)
            k = rng.integers(1, n_levels - 1, size=idx.shape[0])
            s.layer_thickness[idx, k] *= -0.5
        return a, b
    raise ValueError(variant)


def seeds_grid(n_side=15):
    return S.seed_grid(n_side, n_side, (-70.0, 70.0), (-175.0, 175.0))


def seeds_random(n, seed=11):
    return S.uniform_sphere_seeds(n, seed, lat_max=85.0)


@functools.lru_cache(maxsize=None)
def voronoi_mesh(kind: str):
    """general spherical-Voronoi fixtures: 'm8' (jittered icosahedral, 4..8 edges -> 8-wide cell records),
    'm20' (random generators, up to ~12 edges -> 20-wide records)"""
    if kind == "m8":
        return S.voronoi_mesh(S.jittered_icosahedral_points(4, 0.15, 1))
    if kind == "m20":
        return S.voronoi_mesh(S.random_sphere_points(1500, 2))
    raise ValueError(kind)


def voronoi_snapshots(kind: str, n_levels: int):
    m = voronoi_mesh(kind)
    return (S.solid_body_snapshot(m, n_levels, 2.0, tilt=0.3, shear=0.4, bumpy=0.3, w_amp=2e-3, with_attrs=True),
            S.solid_body_snapshot(m, n_levels, 3.0, tilt=0.35, shear=0.2, bumpy=0.25, w_amp=-1e-3, with_attrs=True))


@functools.lru_cache(maxsize=None)
def ocean_mesh(level: int = 4):
    """culled (ocean-only) mesh: continents, a meridional wall with a strait; cellsOnCell / cellsOnVertex hold
    0 where the neighbour was removed"""
    return S.carve_land(mesh(level), S.continents_mask(mesh(level)))


@functools.lru_cache(maxsize=None)
def ocean_snapshots(level: int, n_levels: int):
    m = ocean_mesh(level)
    return (S.solid_body_snapshot(m, n_levels, 2.0, tilt=0.3, shear=0.4, bumpy=0.3, w_amp=2e-3, with_attrs=True),
            S.solid_body_snapshot(m, n_levels, 3.0, tilt=0.35, shear=0.2, bumpy=0.25, w_amp=-1e-3, with_attrs=True))
