"""TEST INFRASTRUCTURE ONLY -- named workloads shared by tests, smoke() and bench.py's baseline legs.

Values restate the reference YAMLs (configs/interm_8m.yaml, configs/interm_117m.yaml: model
section :24-45, default_vars :85-109, dict_out_variables :200-204, var_weights :225-232,
spatial_resolution :76-82).
"""
from __future__ import annotations

DEFAULT_VARS_23 = [
    "land_sea_mask", "orography", "lattitude", "landcover", "2m_temperature", "2m_temperature_max",
    "2m_temperature_min", "temperature_200", "temperature_500", "temperature_850",
    "10m_u_component_of_wind", "u_component_of_wind_200", "u_component_of_wind_500",
    "u_component_of_wind_850", "10m_v_component_of_wind", "v_component_of_wind_200",
    "v_component_of_wind_500", "v_component_of_wind_850", "specific_humidity_200",
    "specific_humidity_500", "specific_humidity_850", "total_precipitation_24hr",
    "volumetric_soil_water_layer_1",
]
OUT_VARS_3 = ["total_precipitation_24hr", "2m_temperature_min", "2m_temperature_max"]
VAR_WEIGHTS = {"2m_temperature": 10, "10m_u_component_of_wind": 1, "10m_v_component_of_wind": 1,
               "total_precipitation_24hr": 1, "2m_temperature_min": 10, "2m_temperature_max": 10}
PRISM_VARS_7 = ["land_sea_mask", "orography", "lattitude", "landcover", "total_precipitation_24hr",
                "2m_temperature_min", "2m_temperature_max"]


def _cfg(default_vars, img, D, depth, dec, heads, res, in_vars=None, out_vars=None, p=2, mag=4, init_img=None):
    in_vars = list(in_vars or default_vars)
    out_vars = list(out_vars or OUT_VARS_3)
    return dict(default_vars=list(default_vars), img_size=tuple(img), init_img_size=tuple(init_img or img),
                in_channels=len(in_vars), out_channels=len(out_vars), patch_size=p, superres_mag=mag, cnn_ratio=4,
                embed_dim=D, depth=depth, decoder_depth=dec, num_heads=heads, mlp_ratio=4.0,
                spatial_resolution=res, in_vars=in_vars, out_vars=out_vars, var_weights=dict(VAR_WEIGHTS))


# tiny: 10 default vars, batch uses a permuted 8-variable subset (exercises var_map indexing)
TINY_DEFAULT = DEFAULT_VARS_23[:7] + ["temperature_850", "total_precipitation_24hr", "volumetric_soil_water_layer_1"]
TINY_IN = ["orography", "2m_temperature_min", "land_sea_mask", "total_precipitation_24hr", "landcover",
           "temperature_850", "lattitude", "2m_temperature_max"]

CASES = {
    # golden fixture case: D=128, 2 heads x 64, depth 2, 8x16 -> 32x64
    "tiny": lambda: _cfg(TINY_DEFAULT, (8, 16), 128, 2, 2, 2, 625.0, in_vars=TINY_IN),
    # tiny with a non-multiple-of-anything grid and the PRISM-like 7-var subset, 12x24 -> 48x96
    "tiny_prism": lambda: _cfg(TINY_DEFAULT, (12, 24), 128, 2, 2, 2, 18.0, in_vars=PRISM_VARS_7),
    # BASELINE.json configs[0]: interm_8m, ERA5 5.625 -> 1.40625, 32x64 -> 128x256
    "8m": lambda: _cfg(DEFAULT_VARS_23, (32, 64), 256, 6, 4, 4, 625.0),
    # BASELINE.json configs[1]: interm_117m, ERA5 1.0 -> 0.25 on the 180x360 -> 720x1440 grid
    "117m": lambda: _cfg(DEFAULT_VARS_23, (180, 360), 1024, 8, 4, 16, 111.0),
    # BASELINE.json configs[2]: interm_1b widths (configs/interm_1b.yaml:39-42: D=3072, 24 heads x 128, depth 8, dec 4),
    # PRISM 7-variable batches at 18 km (:80); "1b_small" keeps the widths on a 16x32 grid with depth 2 / dec 1 so the
    # float64 oracle finishes in seconds, "1b" is the full model on a 64x128 PRISM-like grid
    "1b_small": lambda: _cfg(DEFAULT_VARS_23, (16, 32), 3072, 2, 1, 24, 18.0, in_vars=PRISM_VARS_7),
    "1b": lambda: _cfg(DEFAULT_VARS_23, (64, 128), 3072, 8, 4, 24, 18.0, in_vars=PRISM_VARS_7),
    # BASELINE configs[3]: interm_10b head shape (configs/interm_10b.yaml:39-42: D=8192, 32 heads x 256, depth 11, dec 4);
    # "10b_small" keeps head dim 256 (the shape no 64 / 128 kernel covers) at D=512 so the float64 oracle finishes in
    # seconds; the full widths need the 8-GPU sharded engine
    "10b_small": lambda: _cfg(DEFAULT_VARS_23, (16, 32), 512, 2, 1, 2, 4.0),
    # the full interm_10b model (9.5 B parameters) on the 32x64 ERA5 5.625-degree grid of its YAML's first dataset
    "10b": lambda: _cfg(DEFAULT_VARS_23, (32, 64), 8192, 11, 4, 32, 625.0),
    # interm_10b widths with 2 of the 11 Blocks (2.0 B parameters): every kernel shape of the full model on a 1-2 GPU budget
    "10b_d2": lambda: _cfg(DEFAULT_VARS_23, (32, 64), 8192, 2, 4, 32, 625.0),
    # reduced-grid 117M (same widths, L=4050) for bounded CPU timing
    "117m_90x180": lambda: _cfg(DEFAULT_VARS_23, (90, 180), 1024, 8, 4, 16, 111.0),
}


def get_case(name: str) -> dict:
    return CASES[name]()
