import sys, os, subprocess, tempfile
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from helpers import Golden
from test_host_cli import write_plink
for name, extra in (("uk10k", []), ("groups", ["--group"]), ("gauss", [])):
    g = Golden(name)
    d = tempfile.mkdtemp()
    base = write_plink(g, d, groups_file=(name == "groups"))
    S = ";".join(",".join("%.17g" % s for s in row) for row in g.S)
    args = ["bayesrrcmd_b200/brr_gpu", "--analysis-type", "ppbayes", "--data-file", base + ".bed", "--pheno", base + ".phen", "--chain-length", "60",
            "--burn-in", "0", "--thin", "10", "--seed", "1", "--S", S, "--mcmc-samples", os.path.join(d, "o.csv"), "--no-epsilon"]
    if name == "groups": args += ["--group", base + ".group"]
    r = subprocess.run(args, capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if "mean" in l or "last sweep" in l]
    print(name, "N=%d M=%d" % (g.N, g.M), "|", " | ".join(lines), r.stderr[-200:])
