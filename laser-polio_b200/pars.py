"""Default parameters and component run order: the parameter surface of the reference (pars.py:10-99), kept verbatim
in names, defaults and meaning so existing scripts configure this implementation unchanged."""

from __future__ import annotations

import datetime

import numpy as np

from . import distributions as dist
from .core import PropertySet

__all__ = ["default_pars", "default_run_order"]

default_pars = PropertySet(
    {
        "seed": None,
        # time
        "start_date": datetime.date(2019, 1, 1),
        "dur": 30,
        # population
        "init_pop": [15000, 10000],
        "init_immun": [0.0, 0.0],
        "init_sus_by_age": None,
        "age_pyramid_path": "data/Nigeria_age_pyramid_2024.csv",
        "cbr": [37, 41],
        # disease
        "strain_ids": {"VDPV2": 0, "Sabin2": 1, "nOPV2": 2},
        "strain_r0_scalars": {0: 1.0, 1: 0.25, 2: 0.125},
        "init_prev": [0.0, 0.0],
        "seed_schedule": None,
        "r0": 14,
        "r0_scalars": [0.8, 1.2],
        "seasonal_amplitude": 0.125,
        "seasonal_peak_doy": 180,
        "individual_heterogeneity": True,
        "risk_mult_var": 4.0,
        "corr_risk_inf": 0.8,
        "dur_exp": dist.poisson(lam=3),
        "dur_inf": dist.gamma(shape=4.51, scale=5.32),
        "t_to_paralysis": dist.lognormal(mean=12.5, sigma=3.5),
        "p_paralysis": 1 / 2000,
        # geography
        "shp": None,
        "node_lookup": None,
        "distances": np.array([[0, 100], [100, 0]]),
        # migration
        "node_seeding_dispersion": 1000,
        "node_seeding_zero_inflation": 0.0,
        "migration_method": "radiation",
        "radiation_k_log10": -0.3,
        "gravity_k": 1.0,
        "gravity_k_exponent": 0.0,
        "gravity_a": 1,
        "gravity_b": 1,
        "gravity_c": 2.0,
        "max_migr_frac": 0.1,
        # interventions
        "vx_prob_ri": None,
        "vx_prob_ipv": None,
        "ipv_start_year": 2015,
        "sia_schedule": None,
        "vx_prob_sia": None,
        "missed_frac": 0.0,
        "vx_efficacy": {
            "perfect": 1.0, "bOPV": 0, "f-IPV": 0, "IPV": 0, "IPV + bOPV": 0, "mOPV2": 0.7, "nOPV2": 0.7 * 0.8,
            "nOPV2 + fIPV": 0.7 * 0.8, "topv": 0.5,
        },
        # component step sizes
        "step_size_VitalDynamics_ABM": 7,
        "step_size_DiseaseState_ABM": 1,
        "step_size_RI_ABM": 14,
        "step_size_SIA_ABM": 1,
        "step_size_Transmission_ABM": 1,
        # calibration hooks (unused by the per-tick path)
        "actual_data": None,
        "summary_config": None,
        "verbose": 1,
        "stop_if_no_cases": True,
        "device_init": False,  # extension (not a reference key): draw the per-agent columns on the GPU (popinit.py)
        # extension: newborns get ri_timer = 182 as the reference intends (model.py:1731-1732); False reproduces what it does --
        # its isinstance test on classes never fires, so newborn timers stay at -1 and newborns never receive RI
        "ri_newborn_timer": False,
        # extension: every `compact_every` ticks the device table is compacted -- live agents stably re-sorted by node, the slots
        # of the dead moved out of the swept range (free-slot reuse), host order restored at to_host(); 0 = never (the reference
        # never compacts).  Fused path only.  Changes which slot, hence which random stream, an agent has: DESIGN.md section 3
        "compact_every": 0,
    }
)

default_run_order = ["VitalDynamics_ABM", "DiseaseState_ABM", "RI_ABM", "SIA_ABM", "Transmission_ABM"]
