"""One-off import of the track way-points that define Car(track=...) (data, not code).

Reads the reference's envs/tracks/<name>_track.json (columns X, Y, X_i, Y_i, X_o, Y_o: centre line, inner and
outer border way-points, envs/tracks/tracks.py:14-64) and stores them as float64 arrays in
ilqc_b200/data/track_points.npz.  The splines themselves are re-fitted by ilqc_b200/envs/tracks.py.
Run in the build container only: python tools/import_track_data.py
"""
import json
import os
import numpy as np

REF = "/root/reference/envs/tracks"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ilqc_b200", "data", "track_points.npz")
out = {}
for name in ["simple", "complex", "circle", "line", "bend"]:
    d = json.load(open(os.path.join(REF, name + "_track.json")))
    n = len(d["X"])
    get = (lambda col, i: d[col][str(i)]) if isinstance(d["X"], dict) else (lambda col, i: d[col][i])
    for col in ["X", "Y", "X_i", "Y_i", "X_o", "Y_o"]:
        out[f"{name}_{col}"] = np.array([get(col, i) for i in range(n)], dtype=np.float64)
np.savez_compressed(OUT, **out)
print("wrote", OUT, {k: v.shape for k, v in out.items()})
