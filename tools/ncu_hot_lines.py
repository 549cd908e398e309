"""Hot source lines of one kernel of an Nsight Compute report (run here, no GPU):
    python tools/ncu_hot_lines.py gpurun_out/x.ncu-rep k_lin_schur [top]
Uses `ncu --page source --print-source sass,cuda`: rows with a line number and no address are the per-line aggregates."""
import csv
import subprocess
import sys


def hot_lines(rep: str, kernel: str, top: int = 40) -> str:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kernel}",
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    res, fname, hdr = [], None, None
    for r in rows:
        if r and r[0] == "File Path":
            fname, hdr = r[1].split("/")[-1], None
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr and len(r) == len(hdr) and r[2] == "-":
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            try:
                res.append((int(r[i_s]), int(r[i_i]), fname, r[0], r[1].strip()))
            except ValueError:
                pass
    tot, toti = max(1, sum(o[0] for o in res)), max(1, sum(o[1] for o in res))
    lines = [f"# {rep} {kernel}: {tot} stall samples, {toti} warp instructions"]
    for o in sorted(res, reverse=True)[:top]:
        lines.append(f"{o[0]:6d} {100 * o[0] / tot:5.1f}%  inst {o[1]:8d} {100 * o[1] / toti:4.1f}%  {o[2]}:{o[3]}  {o[4][:110]}")
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    sys.stdout.write(hot_lines(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40))
