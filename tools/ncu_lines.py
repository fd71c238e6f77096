"""Per-source-line instruction and stall-sample shares of one kernel of an ncu report (needs -lineinfo and
--import-source on at capture time).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep <kernel regex> [top N]
"""
import csv
import io
import subprocess
import sys


def main(rep, kern, top=40):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                         capture_output=True, text=True).stdout
    cur_file, hdr, rows = None, None, []
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("Function Name",) or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))
        def num(k):
            try:
                return int(d.get(k, "0") or 0)
            except ValueError:
                return 0
        stalls = {k[6:]: num(k) for k in d if k.startswith("stall_") and "Not Issued" not in k}
        rows.append((cur_file, int(r[0]), r[1].strip()[:100], num("Instructions Executed"), num("# Samples"), stalls))
    ti, ts = sum(x[3] for x in rows) or 1, sum(x[4] for x in rows) or 1
    agg = {}
    for x in rows:
        for k, v in x[5].items():
            agg[k] = agg.get(k, 0) + v
    print(f"kernel {kern}: {ti} warp instructions, {ts} stall samples")
    print("stall reasons:", ", ".join(f"{k} {v * 100 / ts:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 100 / ts >= 1))
    for x in sorted(rows, key=lambda x: -(x[3] / ti + x[4] / ts))[:top]:
        st = sorted(x[5].items(), key=lambda kv: -kv[1])[:2]
        print(f"{x[3] * 100 / ti:5.1f}% inst {x[4] * 100 / ts:5.1f}% smp  {x[0]}:{x[1]:<4} {x[2]}   [{', '.join(f'{k} {v}' for k, v in st if v)}]")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
