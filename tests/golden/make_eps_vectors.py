"""Extract the per-window density vectors the reference plotted for its own example (track A = salmon polygon,
track B = orange polygon of Example/Example_output/single_read_plots_adj/read1.eps) into a small JSON fixture.

Run here (needs /root/reference):  python tests/golden/make_eps_vectors.py
The polygon is polygon(y = c(0, density, last(density), 0), x = c(1, start_index, L, L)) (NanoTel.R:1331-1338);
344.00 pt = density 1.0 (read1.eps axis labels 0.0 at 82.89 pt, 1.0 at 426.89 pt).  This script covers read 1
(30 windows, a single run of relative segments); reads 2-4 are extracted by make_eps_polylines.py, which also follows
the absolute "lineto" anchors the device writes every 100 segments.
"""
import json
import os
import re

SRC = "/root/reference/Example/Example_output/single_read_plots_adj/read1.eps"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_read1_eps_density.json")
COLORS = {"1 0.6471 0": "mismatch", "0.9804 0.5020 0.4471": "exact"}


def main():
    lines = open(SRC).read().splitlines()
    out = {}
    i = 0
    while i < len(lines):
        m = re.match(r"/bg \{ (.+) srgb \} def", lines[i])
        if m and m.group(1) in COLORS:
            name = COLORS[m.group(1)]
            while not lines[i].startswith("np"):
                i += 1
            i += 1
            x0, y0 = map(float, lines[i].split()[:2])
            pts = [(x0, y0)]
            i += 1
            while re.match(r"^-?[\d.]+ -?[\d.]+ l$", lines[i].strip()):
                dx, dy = map(float, lines[i].split()[:2])
                pts.append((pts[-1][0] + dx, pts[-1][1] + dy))
                i += 1
            # vertices: (1,0), (start_1,d_1) .. (start_n,d_n), (L,d_n), (L,0)
            ys = [round((p[1] - y0) / 344.0, 5) for p in pts]
            xs = [round(p[0] - x0, 2) for p in pts]
            out[name] = {"y": ys, "x_pt": xs}
        i += 1
    json.dump({"source": "Example/Example_output/single_read_plots_adj/read1.eps", "pt_per_unit_density": 344.0,
               "tracks": out}, open(OUT, "w"), indent=0)
    for k, v in out.items():
        print(k, len(v["y"]), v["y"][:8])


if __name__ == "__main__":
    main()
