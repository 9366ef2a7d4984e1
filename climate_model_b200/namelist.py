"""User namelist -- same variable names as the reference's namelist.py.

Defaults describe BASELINE.json configs[1]: 1 deg x 1 deg, 32 sigma levels, lat +-84,
elev.1-deg topography, dry dynamical core (physics modules off, coupling fields zero) on
one B200.  Every value can be overridden per run: `Grid(**overrides)` /
`solver.run(**overrides)` take the same names.
"""
import numpy as np

# GRID (namelist.py:24-33)
nb = 1
lon0_deg = 0
lon1_deg = 360
lat0_deg = -84
lat1_deg = 84
dlat_deg = 1.0
dlon_deg = 1.0
nz = 32
pair_top = 10000.

# INITIAL CONDITIONS (namelist.py:37-58)
gaussian_dlon = np.pi / 10
gaussian_dlat = np.pi / 10
uwind_0 = 0
vwind_0 = 0
UWIND_gaussian_pert = 10
UWIND_random_pert = 0
VWIND_gaussian_pert = 10
VWIND_random_pert = 0
COLP_gaussian_pert = 0
COLP_random_pert = 0
POTT_gaussian_pert = 0
POTT_random_pert = 0.0
QV_gaussian_pert = 0.0
QV_random_pert = 0.0

# DYNAMICS SWITCHES (namelist.py:62-88); the B200 dyn core implements the all-on
# configuration; the *_vert_turb / radiation terms act on zero coupling fields
i_COLP_main_switch = 1
i_UVFLX_main_switch = 1
i_UVFLX_hor_adv = 1
i_UVFLX_vert_adv = 1
i_UVFLX_vert_turb = 1
i_UVFLX_coriolis = 1
i_UVFLX_num_dif = 1
i_UVFLX_pre_grad = 1
i_POTT_main_switch = 1
i_POTT_hor_adv = 1
i_POTT_vert_adv = 1
i_POTT_vert_turb = 1
i_POTT_num_dif = 1
i_POTT_radiation = 1
i_POTT_microphys = 1
i_moist_main_switch = 1
i_moist_hor_adv = 1
i_moist_vert_adv = 1
i_moist_vert_turb = 1
i_moist_num_dif = 1
i_moist_microphys = 1

# TOPOGRAPHY (namelist.py:94-95)
i_use_topo = 1
n_topo_smooth = 10

# PHYSICS MODULES.  The turbulence module is built (turb_main.py; i_turbulence = 1 makes the
# grid evaluate the coupled terms); the surface, radiation and microphysics modules are out
# of scope of this package (always off): their fields enter through the coupling inputs
i_surface_scheme = 0
i_radiation = 0
i_microphysics = 0
i_turbulence = 0

# IO / RUN CONTROL (namelist.py:146-240)
nth_ts_print_diag = 50
i_out_nth_hour = 12.0
i_sim_n_days = 1.0
output_path = '../output'
i_load_from_restart = 0
i_save_to_restart = 0
i_restart_nth_day = 1.0
i_load_from_IC = 0
i_time_stepping = 'MATSUNO'
CFL = 0.7
working_precision = 'float64'
# 1: numba CPU and 2: numba GPU exist only in the reference; 3: B200 CUDA dyn core
i_comp_mode = 3
i_sync_context = 0

# DIFFUSION (namelist.py:286-318)
POTT_dif_coef = 1E-5
COLP_dif_coef = 0
moist_dif_coef = POTT_dif_coef


def UVFLX_dif_coef_for(dlat_deg):
    """UVFLX_dif_coef as the reference chooses it from the resolution (namelist.py:296-316)"""
    table = {10: 1.5, 8: 1.5, 6: 1.8, 5: 2, 4: 2.5, 3: 3.3, 2: 5, 1.5: 7.5}
    coef = 0
    if dlat_deg in table:
        coef = table[dlat_deg]
    elif dlat_deg <= 1:
        coef = 10
    return coef * 2.0


UVFLX_dif_coef = UVFLX_dif_coef_for(dlat_deg)
