"""Extract the density polygons the reference plotted for reads 2-4 of its own example into a JSON fixture
(track A = salmon polygon, track B = orange polygon of Example/Example_output/single_read_plots_adj/read<k>.eps).

Run here (needs /root/reference):  python tests/golden/make_eps_polylines.py
The polygon is polygon(y = c(0, density, last(density), 0), x = c(1, start_index, L, L)) (NanoTel.R:1331-1338).
R's PostScript device writes relative segments rounded to 0.01 pt and re-anchors with an absolute "lineto" every 100
segments; one vertex per window, so the test can compare window by window (to the device resolution).  Coordinates are kept in PostScript points relative to the
polygon's first vertex (position 1, density 0); 344 pt = density 1.0 (axis labels 0.0 at 82.89 pt, 1.0 at 426.89 pt).
"""
import json
import os
import re

DIR = "/root/reference/Example/Example_output/single_read_plots_adj"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_reads234_eps_polylines.json")
COLORS = {"1 0.6471 0": "mismatch", "0.9804 0.5020 0.4471": "exact"}


def polygons(path):
    lines = open(path).read().splitlines()
    out = {}
    i = 0
    while i < len(lines):
        m = re.match(r"/bg \{ (.+) srgb \} def", lines[i])
        if m and m.group(1) in COLORS and COLORS[m.group(1)] not in out:
            while not lines[i].startswith("np"):
                i += 1
            i += 1
            x0, y0 = map(float, lines[i].split()[:2])
            pts = [(x0, y0)]
            i += 1
            while True:                          # relative "dx dy l"; every 100 segments an absolute "x y lineto"
                ln = lines[i].strip()
                if re.match(r"^-?[\d.]+ -?[\d.]+ l$", ln):
                    dx, dy = map(float, ln.split()[:2])
                    pts.append((pts[-1][0] + dx, pts[-1][1] + dy))
                elif re.match(r"^-?[\d.]+ -?[\d.]+ lineto$", ln):
                    pts.append(tuple(map(float, ln.split()[:2])))
                else:
                    break
                i += 1
            out[COLORS[m.group(1)]] = {"x_pt": [round(p[0] - x0, 2) for p in pts], "y_pt": [round(p[1] - y0, 2) for p in pts]}
        i += 1
    return out


def main():
    res = {"source": "Example/Example_output/single_read_plots_adj/read{2,3,4}.eps", "pt_per_unit_density": 344.0, "reads": {}}
    for k in (2, 3, 4):
        res["reads"][str(k)] = polygons(os.path.join(DIR, "read%d.eps" % k))
        for name, v in res["reads"][str(k)].items():
            print(k, name, len(v["x_pt"]), "vertices, width", v["x_pt"][-1], "pt")
    json.dump(res, open(OUT, "w"))


if __name__ == "__main__":
    main()
