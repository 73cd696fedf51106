"""Extract two more known-answer artefacts of the reference's own example run into a small JSON fixture:

* Example/Example_output/log/run.log:20-46 -- R's summary() (Min, 1st Qu., Median, Mean, 3rd Qu., Max) of the read
  lengths, the telomere lengths (track A) and the telomere lengths with one mismatch (track B) (NanoTel.R:2396-2427);
* the subtitle of every single_read_plots_adj/read<k>.eps ("Read length: .. , Telomere length: .. , Telomere length
  with mismatches: ..", NanoTel.R:1876-1918 -> plot_single_telo_with_gray_area :1271-1410).

Run here (needs /root/reference):  python tests/golden/make_log_and_subtitles.py
"""
import json
import os
import re

REF = "/root/reference/Example/Example_output"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_log_and_subtitles.json")


def main():
    log = open(os.path.join(REF, "log", "run.log")).read().splitlines()
    summaries = []
    for i, line in enumerate(log):
        if line.split() == ["Min.", "1st", "Qu.", "Median", "Mean", "3rd", "Qu.", "Max."]:
            summaries.append([float(x) for x in log[i + 1].split()])
    assert len(summaries) == 4, summaries        # sample reads, telomeric reads, telomere length, ... with mismatch
    out = {"source": "Example/Example_output/log/run.log:20-46 and single_read_plots_adj/read*.eps subtitles",
           "summary_names": ["Min.", "1st Qu.", "Median", "Mean", "3rd Qu.", "Max."],
           "read_length": summaries[0], "telomeric_read_length": summaries[1],
           "telomere_length": summaries[2], "telomere_length_mismatch": summaries[3], "subtitles": {}}
    for k in (1, 2, 3, 4):
        txt = open(os.path.join(REF, "single_read_plots_adj", "read%d.eps" % k)).read()
        # the PostScript device splits the string at kerning pairs: "(Read length: 2981 , T) ... (elomere length: 2976 , T) ..."
        flat = "".join(re.findall(r"\(([^()]*)\)", txt))
        m = re.search(r"Read length: (\d+) , Telomere length: (\d+) , Telomere length with mismatches: (\d+)", flat)
        assert m, k
        out["subtitles"][str(k)] = {"read_length": int(m.group(1)), "telomere_length": int(m.group(2)),
                                    "telomere_length_mismatch": int(m.group(3))}
    json.dump(out, open(OUT, "w"), indent=1)
    print(OUT, out)


if __name__ == "__main__":
    main()
