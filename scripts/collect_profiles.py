"""Turns what scripts/round_profile.sh left in gpurun_out/ into the tracked summaries under profiles/:
    python scripts/collect_profiles.py r02        # writes profiles/r02_*.{json,csv,txt} and roofline_traffic.json
Needs ncu (reads the .ncu-rep files with `ncu -i`).  COLLECT_OUT=<dir> writes there instead: the reports of a round exceed
what gpurun copies back (64 MiB), so scripts/r2_final2.sh summarises them on the GPU box and deletes them."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.environ.get("COLLECT_OUT") or os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEEP = ("dram__bytes", "gpu__dram_throughput", "gpu__time_duration", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate",
        "launch__", "sm__inst_executed", "sm__pipe_", "sm__warps_active", "smsp__average_warps_issue_stalled",
        "smsp__issue_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed", "sm__throughput",
        "smsp__cycles_active.avg")


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def clean_csv(path):
    return "".join(l for l in open(path) if not l.startswith("=="))


def main(prefix):
    out = lambda name: os.path.join(P, "%s_%s" % (prefix, name))
    open(out("bench_line.json"), "w").write(open(os.path.join(G, "bench_full.json")).read().strip().splitlines()[-1] + "\n")
    open(out("launches_bench_default.csv"), "w").write("".join(clean_csv(os.path.join(G, "launches.csv")).splitlines(True)[:41]))
    dram = clean_csv(os.path.join(G, "dram_1m.csv"))
    open(out("dram_bytes_1m_launch.csv"), "w").write(dram)
    vals = {r[12]: int(r[14]) for r in csv.reader(io.StringIO(dram)) if len(r) > 14 and r[14].isdigit()}
    rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
    json.dump({"kernel": "gmapdp_dp_kernel<0> (single gaps: full fills)", "dram_bytes_per_launch": rd + wr,
               "source": "profiles/%s_dram_bytes_1m_launch.csv: ncu --kernel-name-base mangled -k regex:ILi0E --metrics "
                         "dram__bytes_read.sum,dram__bytes_write.sum on `python bench.py --no-cpu-baseline --chain-problems 0 "
                         "--steps 1 --warmup 0` (1M boxes; the single-gap kernel over the 200k single-gap boxes): "
                         "%.2f GB read + %.2f GB written" % (prefix, rd / 1e9, wr / 1e9)},
              open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    for rep, name in (("prof_full", "single_kernel_100k"), ("prof_genome", "genome_kernel_100k"), ("prof_cdna", "cdna_kernel_100k"),
                      ("prof_end", "end_kernel_100k"), ("prof_mepass", "maxent_kernel_100k"),
                      ("prof_small2_genome", "small_genome_kernel_in_step"), ("prof_small2_end", "small_end_kernel_in_step")):
        path = os.path.join(G, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
        h, u, v = rows[0], rows[1], rows[2]
        with open(out(name + "_metrics.csv"), "w") as f:
            f.write("kernel,%s\nmetric,unit,value\n" % v[h.index("Kernel Name")])
            for i, n in enumerate(h):
                if n.startswith(KEEP):
                    f.write("%s,%s,%s\n" % (n, u[i], v[i]))
        src = os.path.join(G, rep + "_src.csv")
        open(src, "w").write(ncu(["-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]))
        hot = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), src, "45"],
                             stdout=subprocess.PIPE, text=True).stdout
        open(out(name + "_hot_lines.txt"), "w").write(hot)
    print("wrote profiles/%s_*" % prefix)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "rXX")
