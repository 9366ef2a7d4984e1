"""Initial conditions and sigma levels (reference: io_initial_conditions.py:31-287,
mic_main.py:63-71), vectorised numpy instead of the reference's Python loops.

Inputs are the reference's two data assets, stored as arrays in data/ic_data.npz
(tools/import_reference_data.py).  Host-side set-up code, not on the timed path.
"""
import os

import numpy as np
from scipy.interpolate import RectBivariateSpline, interp1d

from . import namelist as nl
from .io_constants import con_kappa, wp

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'ic_data.npz')
_cache = {}


def _data():
    if 'd' not in _cache:
        _cache['d'] = dict(np.load(_DATA))
    return _cache['d']


def _bilinear(x, y, z, xnew, ynew):
    """bilinear interpolation with edge clamping: stands in for scipy's removed
    interp2d(kind='linear') that io_initial_conditions.py:191-193 called"""
    ix, iy = np.argsort(x), np.argsort(y)
    xs, ys = x[ix], y[iy]
    spl = RectBivariateSpline(xs, ys, z[iy, :][:, ix].T.astype(np.float64), kx=1, ky=1, s=0)
    xc = np.clip(xnew, xs[0], xs[-1])
    yc = np.clip(ynew, ys[0], ys[-1])
    jx, jy = np.argsort(xc), np.argsort(yc)
    out_sorted = spl(xc[jx], yc[jy])
    out = np.empty_like(out_sorted)
    out[np.ix_(jx, jy)] = out_sorted
    return out                                   # (len(xnew), len(ynew))


def load_topo(GR, HSURF, n_topo_smooth=None):
    """io_initial_conditions.py:185-213"""
    n_topo_smooth = nl.n_topo_smooth if n_topo_smooth is None else n_topo_smooth
    d = _data()
    ii, jj, nb = GR.ii, GR.jj, GR.nb
    HSURF[ii, jj, 0] = _bilinear(d['elev_lon'], d['elev_lat'], d['elev'],
                                 GR.lon_deg[ii, nb + 1, 0].squeeze(),
                                 GR.lat_deg[nb + 1, jj, 0].squeeze())
    HSURF = GR.exchange_BC(HSURF)
    HSURF[HSURF < 0] = 0
    HSURF = GR.exchange_BC(HSURF)
    tau_smooth_min, tau_smooth_max = 0.05, 0.15
    tau = np.full((GR.nx + 2 * nb, GR.ny + 2 * nb, 1), np.nan)
    j = np.arange(nb, nb + GR.ny)
    tau[:, j, 0] = tau_smooth_min + np.sin(GR.lat_rad[5, j, 0]) ** 2 * (
        tau_smooth_max - tau_smooth_min)
    for _ in range(n_topo_smooth):
        HSURF[ii, jj] = HSURF[ii, jj] + (tau[ii, jj] * (
            HSURF[ii - 1, jj] + HSURF[ii + 1, jj] + HSURF[ii, jj - 1] + HSURF[ii, jj + 1]
            - 4 * HSURF[ii, jj]))
        HSURF = GR.exchange_BC(HSURF)
    return HSURF


def set_up_sigma_levels(GR, pair_top=None):
    """io_initial_conditions.py:253-287 (quadratic-in-height interfaces; the mean surface
    height always comes from the smoothed topography, as in the reference)"""
    pair_top = nl.pair_top if pair_top is None else pair_top
    HSURF = np.full((GR.nx + 2 * GR.nb, GR.ny + 2 * GR.nb, 1), np.nan, dtype=wp)
    HSURF = load_topo(GR, HSURF)
    profile = _data()['profile']
    zsurf_test = np.mean(HSURF[GR.ii, GR.jj])
    top_ind = np.argwhere(profile[:, 2] >= pair_top).squeeze()[-1]
    ztop_test = profile[top_ind, 0] + (profile[top_ind, 2] - pair_top) / (
        profile[top_ind, 4] * profile[top_ind, 1])
    ks = np.arange(0, GR.nzs)
    z_vb_test = np.zeros(GR.nzs, dtype=wp)
    p_vb_test = np.zeros(GR.nzs, dtype=wp)
    z_vb_test[0] = ztop_test
    z_vb_test[ks] = zsurf_test + (ztop_test - zsurf_test) * (1 - ks / GR.nz) ** (2)
    rho_vb_test = np.interp(z_vb_test, profile[:, 0], profile[:, 4])
    g_vb_test = np.interp(z_vb_test, profile[:, 0], profile[:, 1])
    p_vb_test[0] = pair_top
    for k in range(1, GR.nzs):
        p_vb_test[k] = p_vb_test[k - 1] + rho_vb_test[k] * g_vb_test[k] * (
            z_vb_test[k - 1] - z_vb_test[k])
    GR.sigma_vb[:] = (p_vb_test - pair_top) / (p_vb_test[-1] - pair_top)
    GR.dsigma[:] = np.diff(GR.sigma_vb)


def _gaussian2D(GR, FIELD, pert, lon0_rad, lat0_rad, lonSig_rad, latSig_rad):
    """io_initial_conditions.py:224-249"""
    dimx, dimy = FIELD.shape
    if dimy == GR.nys + 2 * GR.nb:
        sel = (GR.ii, GR.jjs)
        lat, lon = GR.lat_js_rad[GR.ii, GR.jjs, 0], GR.lon_js_rad[GR.ii, GR.jjs, 0]
    elif dimx == GR.nxs + 2 * GR.nb:
        sel = (GR.iis, GR.jj)
        lat, lon = GR.lat_is_rad[GR.iis, GR.jj, 0], GR.lon_is_rad[GR.iis, GR.jj, 0]
    else:
        sel = (GR.ii, GR.jj)
        lat, lon = GR.lat_rad[GR.ii, GR.jj, 0], GR.lon_rad[GR.ii, GR.jj, 0]
    perturb = pert * np.exp(- np.power(lon - lon0_rad, 2) / (2 * lonSig_rad ** 2)
                            - np.power(lat - lat0_rad, 2) / (2 * latSig_rad ** 2))
    FIELD[sel] = FIELD[sel] + perturb.squeeze()
    return FIELD


def _random2D(FIELD, pert):
    """io_initial_conditions.py:217-220 (draws even when pert == 0, like the reference)"""
    FIELD[:] = FIELD[:] + pert * np.random.rand(FIELD.shape[0], FIELD.shape[1])
    return FIELD


def _pvt_factor(GR, COLP, PVTF, PVTFVB, pair_top):
    """io_initial_conditions.py:133-150"""
    ii, jj = GR.ii, GR.jj
    PAIRVB = np.full((GR.nx + 2 * GR.nb, GR.ny + 2 * GR.nb, GR.nzs), np.nan, dtype=wp)
    for ks in range(GR.nzs):
        PAIRVB[ii, jj, ks] = pair_top + (GR.sigma_vb[0, 0, ks] * COLP[ii, jj, 0])
    PVTFVB[:] = np.power(PAIRVB / 100000., con_kappa)
    for k in range(GR.nz):
        PVTF[:, :, k][ii, jj] = 1 / (1 + con_kappa) * (
            PVTFVB[:, :, k + 1][ii, jj] * PAIRVB[:, :, k + 1][ii, jj] -
            PVTFVB[:, :, k][ii, jj] * PAIRVB[:, :, k][ii, jj]) / (
            PAIRVB[:, :, k + 1][ii, jj] - PAIRVB[:, :, k][ii, jj])
    return PVTF, PVTFVB


def calc_specific_humidity(T, RH, p):
    """misc_meteo_utilities.py:18-35"""
    T = T - wp(273.15)
    f_p = wp(1.0016) + wp(3.15E-6) * p / wp(100) - wp(0.074) / (p / wp(100))
    esw = wp(100) * f_p * wp(6.112) * np.exp((wp(17.62) * T) / (wp(243.12) + T))
    return RH / wp(100) / p * wp(0.622) * esw


def initialize_fields(GR, host, **pert):
    """Build the initial state in the host field dict `host` (reference layout):
    io_initial_conditions.py:31-131 followed by Microphysics.initial_conditions
    (mic_main.py:63-71, QV at 60 % relative humidity).  `pert` overrides the namelist's
    *_gaussian_pert / *_random_pert / i_use_topo values."""
    P = {k: getattr(nl, k) for k in (
        'uwind_0', 'vwind_0', 'UWIND_gaussian_pert', 'UWIND_random_pert', 'VWIND_gaussian_pert',
        'VWIND_random_pert', 'COLP_gaussian_pert', 'COLP_random_pert', 'POTT_gaussian_pert',
        'POTT_random_pert', 'QV_gaussian_pert', 'QV_random_pert', 'gaussian_dlon',
        'gaussian_dlat', 'i_use_topo')}
    for k, v in pert.items():
        if k not in P:
            raise KeyError('unknown initial-condition parameter %r' % k)
        P[k] = v
    pair_top = wp(GR.pair_top)
    ii, jj, iis, jjs = GR.ii, GR.jj, GR.iis, GR.jjs
    (POTTVB, WWIND, HSURF, COLP, PVTF, PVTFVB, POTT, UWIND, VWIND, QV, QC) = (
        host[n] for n in ('POTTVB', 'WWIND', 'HSURF', 'COLP', 'PVTF', 'PVTFVB', 'POTT', 'UWIND',
                          'VWIND', 'QV', 'QC'))
    np.random.seed(seed=3)
    POTTVB[:] = 0
    WWIND[:] = 0
    if P['i_use_topo']:
        HSURF = load_topo(GR, HSURF)
    else:
        HSURF[:] = 0.

    # set_up_profile, io_initial_conditions.py:154-181
    profile = _data()['profile']
    PSURF = np.full_like(HSURF, np.nan)
    PSURF[ii, jj, 0] = np.interp(HSURF[ii, jj, 0], profile[:, 0], profile[:, 2])
    COLP[ii, jj, 0] = PSURF[ii, jj, 0] - pair_top
    PVTF, PVTFVB = _pvt_factor(GR, COLP, PVTF, PVTFVB, pair_top)
    PAIR = np.full_like(POTT, np.nan)
    TAIR = np.full_like(POTT, np.nan)
    PAIR[ii, jj] = 100000. * np.power(PVTF[ii, jj], 1 / con_kappa)
    TAIR[ii, jj] = interp1d(profile[:, 2], profile[:, 3])(PAIR[ii, jj])
    POTT[ii, jj] = TAIR[ii, jj] * np.power(100000. / PAIR[ii, jj], con_kappa)

    QV[ii, jj, :] = 0.
    QC[ii, jj, :] = 0.
    g = (np.pi * 3 / 4, 0, P['gaussian_dlon'], P['gaussian_dlat'])
    COLP[:, :, 0] = _gaussian2D(GR, COLP[:, :, 0], P['COLP_gaussian_pert'], *g)
    COLP[:, :, 0] = _random2D(COLP[:, :, 0], P['COLP_random_pert'])
    for k in range(GR.nz):
        UWIND[iis, jj, k] = P['uwind_0']
        UWIND[:, :, k] = _gaussian2D(GR, UWIND[:, :, k], P['UWIND_gaussian_pert'], *g) * (
            1 - (k + 1) / GR.nz) ** (1 / 2)
        UWIND[:, :, k] = _random2D(UWIND[:, :, k], P['UWIND_random_pert'])
        VWIND[:, :, k][ii, jjs] = P['vwind_0']
        VWIND[:, :, k] = _gaussian2D(GR, VWIND[:, :, k], P['VWIND_gaussian_pert'], *g) * (
            1 - (k + 1) / GR.nz) ** (1 / 2)
        VWIND[:, :, k] = _random2D(VWIND[:, :, k], P['VWIND_random_pert'])
        POTT[:, :, k] = _gaussian2D(GR, POTT[:, :, k], P['POTT_gaussian_pert'], *g)
        POTT[:, :, k] = _random2D(POTT[:, :, k], P['POTT_random_pert'])
        QV[:, :, k] = _gaussian2D(GR, QV[:, :, k], P['QV_gaussian_pert'], *g)
        QV[:, :, k] = _random2D(QV[:, :, k], P['QV_random_pert'])
    for n, a in (('COLP', COLP), ('UWIND', UWIND), ('VWIND', VWIND), ('POTT', POTT),
                 ('QV', QV), ('QC', QC)):
        host[n] = GR.exchange_BC(a)
    host['HSURF'] = HSURF

    # diagnose_fields_init + diagnose_secondary_fields (io_initial_conditions.py:294-403)
    # only to obtain TAIR / PAIR of the PERTURBED state for the moisture profile
    PVTF, PVTFVB = _pvt_factor(GR, host['COLP'], PVTF, PVTFVB, pair_top)
    TAIR[ii, jj] = host['POTT'][ii, jj] * PVTF[ii, jj]
    PAIR[ii, jj] = 100000 * np.power(PVTF[ii, jj], 1 / con_kappa)
    # Microphysics.initial_conditions, mic_main.py:63-71 (RH_init = 60 %)
    QV = calc_specific_humidity(TAIR, wp(60), PAIR)
    host['QV'] = GR.exchange_BC(QV)
    host['PVTF'], host['PVTFVB'] = PVTF, PVTFVB
    return host
