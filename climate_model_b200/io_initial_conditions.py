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


# ---------------------------------------------------------------------------------------
# band-local initial state (latitude-band runs, SURVEY.md 8e / BASELINE configs[4])
# ---------------------------------------------------------------------------------------
def _bc_window(GR, F, ja, jb, stgx, stgy):
    """misc_boundaries.py:22-42 / main_grid.py:319-358 on a ROW WINDOW [ja, jb] (global rows) of
    a reference-layout array: periodic images in x on every row, wall rows if the window holds
    them"""
    nx, ny, nys = int(GR.nx), int(GR.ny), int(GR.nys)
    if stgx:
        F[0] = F[nx]            # = nxs - 1
        F[nx + 1] = F[1]        # = nxs        (index nxs + 1 stays unset, like the reference)
    else:
        F[0] = F[nx]
        F[nx + 1] = F[1]
    if stgy:
        for j in (0, 1, nys, nys + 1):
            if ja <= j <= jb:
                F[:, j - ja] = 0.
    else:
        if ja <= 0 and jb >= 1:
            F[:, 0 - ja] = F[:, 1 - ja]
        if ja <= ny and jb >= ny + 1:
            F[:, ny + 1 - ja] = F[:, ny - ja]
    return F


def initialize_fields_band(GR, host, rows_of, diagnostics=True, **pert):
    """initialize_fields for ONE latitude band: `host[n]` holds the rows rows_of(n) = (ja, jb)
    of field n only, shape (fnx, jb-ja+1, nk).  Everything 2-D (topography, its smoothing, the
    surface pressure, COLP and its perturbations, the random-number draws) is built for the
    whole grid -- a 2-D field is small at any resolution -- and every 3-D field directly on the
    band's rows: at 0.1 deg x 96 levels a rank of an 8-GPU run builds 217 of 1682 rows
    (0.6 GB per field instead of 4.6 GB).  Same expressions, element by element, as
    initialize_fields (tests/test_ic_builder.py compares the two)."""
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
    nx, ny, nz, nzs, nb = int(GR.nx), int(GR.ny), int(GR.nz), int(GR.nzs), int(GR.nb)
    ii, jj = GR.ii, GR.jj
    np.random.seed(seed=3)
    shape2 = (nx + 2 * nb, ny + 2 * nb, 1)
    HSURF = np.full(shape2, np.nan, dtype=wp)
    if P['i_use_topo']:
        HSURF = load_topo(GR, HSURF)
    else:
        HSURF[:] = 0.
    profile = _data()['profile']
    PSURF = np.full_like(HSURF, np.nan)
    PSURF[ii, jj, 0] = np.interp(HSURF[ii, jj, 0], profile[:, 0], profile[:, 2])
    COLP0 = np.full_like(HSURF, np.nan)                     # before its perturbations
    COLP0[ii, jj, 0] = PSURF[ii, jj, 0] - pair_top
    g = (np.pi * 3 / 4, 0, P['gaussian_dlon'], P['gaussian_dlat'])
    COLP = COLP0.copy()
    COLP[:, :, 0] = _gaussian2D(GR, COLP[:, :, 0], P['COLP_gaussian_pert'], *g)
    COLP[:, :, 0] = _random2D(COLP[:, :, 0], P['COLP_random_pert'])
    COLP = GR.exchange_BC(COLP)
    any_random = any(P[k] != 0 for k in ('UWIND_random_pert', 'VWIND_random_pert',
                                         'POTT_random_pert', 'QV_random_pert'))

    def window(n):
        ja, jb = rows_of(n)
        return ja, jb, max(1, ja), min(ny if n != 'VWIND' else ny + 1, jb)   # + interior rows

    def pvt(colp_rows):
        """_pvt_factor on rows: colp_rows (nx, rows) -> PVTF (nx, rows, nz), PVTFVB (.., nzs)"""
        sig = np.asarray(GR.sigma_vb).reshape(-1)
        PAIRVB = pair_top + sig[None, None, :] * colp_rows[:, :, None]
        PVTFVB = np.power(PAIRVB / 100000., con_kappa)
        PVTF = 1 / (1 + con_kappa) * (PVTFVB[:, :, 1:] * PAIRVB[:, :, 1:] -
                                      PVTFVB[:, :, :-1] * PAIRVB[:, :, :-1]) / (
            PAIRVB[:, :, 1:] - PAIRVB[:, :, :-1])
        return PVTF, PVTFVB

    def gauss(lon, lat, amp):
        return amp * np.exp(- np.power(lon - g[0], 2) / (2 * g[2] ** 2)
                            - np.power(lat - g[1], 2) / (2 * g[3] ** 2))

    i = np.arange(nb, nx + nb)
    i_s = np.arange(nb, nx + 1 + nb)
    kfac = np.array([(1 - (k + 1) / nz) ** (1 / 2) for k in range(nz)])   # as the scalar loop
    # random draws, in the order initialize_fields makes them (whole-grid 2-D arrays per
    # level and field): only needed when a random perturbation is switched on
    rnd = {}
    if any_random:
        for k in range(nz):
            for n, shp in (('UWIND', (nx + 3, ny + 2)), ('VWIND', (nx + 2, ny + 3)),
                           ('POTT', (nx + 2, ny + 2)), ('QV', (nx + 2, ny + 2))):
                ja, jb = rows_of(n)
                rnd[(n, k)] = np.random.rand(*shp)[:, ja:jb + 1]

    # ---- mass fields: POTT, QV, QC, PVTF, PVTFVB -------------------------------------------
    ja, jb, j0, j1 = window('POTT')
    j = np.arange(j0, j1 + 1)
    PVTF0, _ = pvt(COLP0[np.ix_(i, j)][:, :, 0])
    PAIR = 100000. * np.power(PVTF0, 1 / con_kappa)
    TAIR = interp1d(profile[:, 2], profile[:, 3])(PAIR)
    POTT_in = TAIR * np.power(100000. / PAIR, con_kappa)
    lon = GR.lon_rad[np.ix_(i, j)][:, :, 0]
    lat = GR.lat_rad[np.ix_(i, j)][:, :, 0]
    POTT_in = POTT_in + gauss(lon, lat, P['POTT_gaussian_pert'])[:, :, None]
    POTT = np.full((nx + 2, jb - ja + 1, nz), np.nan, dtype=wp)
    POTT[np.ix_(i, j - ja)] = POTT_in
    if any_random:
        for k in range(nz):
            POTT[:, :, k] = POTT[:, :, k] + P['POTT_random_pert'] * rnd[('POTT', k)]
    host['POTT'][...] = _bc_window(GR, POTT, ja, jb, 0, 0)
    PVTF1, PVTFVB1 = pvt(COLP[np.ix_(i, j)][:, :, 0])
    TAIR1 = POTT[np.ix_(i, j - ja)] * PVTF1
    PAIR1 = 100000 * np.power(PVTF1, 1 / con_kappa)
    QV = np.full_like(POTT, np.nan)
    QV[np.ix_(i, j - ja)] = calc_specific_humidity(TAIR1, wp(60), PAIR1)
    host['QV'][...] = _bc_window(GR, QV, ja, jb, 0, 0)
    QC = np.full_like(POTT, np.nan)
    QC[np.ix_(i, j - ja)] = 0.
    host['QC'][...] = _bc_window(GR, QC, ja, jb, 0, 0)
    if diagnostics:
        # PVTF / PVTFVB of the initial state and the zeros of POTTVB / WWIND
        # (io_initial_conditions.py:45-46): the device recomputes / already holds them, so a
        # band-local run skips these four host arrays (0.6 GB each at 0.1 deg x 96 levels)
        for n, a in (('PVTF', PVTF1), ('PVTFVB', PVTFVB1)):
            out = host[n]
            out[...] = np.nan
            out[np.ix_(i, j - ja)] = a
        host['POTTVB'][...] = 0.
        host['WWIND'][...] = 0.
    del PVTF0, PVTF1, PVTFVB1, PAIR, PAIR1, TAIR, TAIR1, POTT_in, QV, QC
    for n in ('COLP', 'HSURF'):
        ja2, jb2 = rows_of(n)
        host[n][...] = (COLP if n == 'COLP' else HSURF)[:, ja2:jb2 + 1]

    # ---- UWIND (x-staggered) -------------------------------------------------------------------
    ja, jb, j0, j1 = window('UWIND')
    j = np.arange(j0, j1 + 1)
    lon = GR.lon_is_rad[np.ix_(i_s, j)][:, :, 0]
    lat = GR.lat_is_rad[np.ix_(i_s, j)][:, :, 0]
    base = P['uwind_0'] + gauss(lon, lat, P['UWIND_gaussian_pert'])
    U = np.full((nx + 3, jb - ja + 1, nz), np.nan, dtype=wp)
    U[np.ix_(i_s, j - ja)] = base[:, :, None] * kfac[None, None, :]
    if any_random:
        for k in range(nz):
            U[:, :, k] = U[:, :, k] + P['UWIND_random_pert'] * rnd[('UWIND', k)]
    host['UWIND'][...] = _bc_window(GR, U, ja, jb, 1, 0)

    # ---- VWIND (y-staggered) -------------------------------------------------------------------
    ja, jb, j0, j1 = window('VWIND')
    j = np.arange(j0, j1 + 1)
    lon = GR.lon_js_rad[np.ix_(i, j)][:, :, 0]
    lat = GR.lat_js_rad[np.ix_(i, j)][:, :, 0]
    base = P['vwind_0'] + gauss(lon, lat, P['VWIND_gaussian_pert'])
    V = np.full((nx + 2, jb - ja + 1, nz), np.nan, dtype=wp)
    V[np.ix_(i, j - ja)] = base[:, :, None] * kfac[None, None, :]
    if any_random:
        for k in range(nz):
            V[:, :, k] = V[:, :, k] + P['VWIND_random_pert'] * rnd[('VWIND', k)]
    host['VWIND'][...] = _bc_window(GR, V, ja, jb, 0, 1)
    return host
