"""One-off import of the track way-points that define Car(track=...) (data, not code).

Reads the reference's envs/tracks/<name>_track.json (columns X, Y, X_i, Y_i, X_o, Y_o: centre line, inner and
outer border way-points) EXACTLY as the reference does, through pandas.read_json (envs/tracks/tracks.py:38-44;
pandas' default JSON float parser differs from Python's json module in the last bit for ~20 % of the values, and
the reference's numbers are the pandas ones), and stores them as float64 arrays in
ilqc_b200/data/track_points.npz.  The splines themselves are re-fitted by ilqc_b200/envs/tracks.py.
Also prints the Pacejka constants of envs/bicycle_model.json as pandas parses them (car.py:104-106); they are
written out literally in ilqc_b200/envs/car.py.
Run in the build container only: python tools/import_track_data.py
"""
import os

import numpy as np
import pandas as pd

REF = "/root/reference/envs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ilqc_b200", "data", "track_points.npz")
out = {}
for name in ["simple", "complex", "circle", "line", "bend"]:
    td = pd.read_json(os.path.join(REF, "tracks", name + "_track.json"))
    for col in ["X", "Y", "X_i", "Y_i", "X_o", "Y_o"]:
        out[f"{name}_{col}"] = td[col].to_numpy(dtype=float)
np.savez_compressed(OUT, **out)
print("wrote", OUT)
s = pd.read_json(os.path.join(REF, "bicycle_model.json"), typ="series")
print({k: float(s[k]) for k in ["Cm1", "Cm2", "Cr0", "Cr2", "Br", "Cr", "Dr", "Bf", "Cf", "Df", "m", "Iz", "lf", "lr",
                               "car_l", "car_w"]})
