"""One-shot GPU session for the end of a round (run under gpurun): the new-kernel parity tests, kernel timings, the
bench matrix (correlation kernel sequence x pairs in flight), the final default bench line, an ncu capture of the
correlation kernel and the step profile.  Every stage has its own timeout and writes into gpurun_out/ as it goes, most
important first, so a cut-off call still leaves the earlier results.

    python scripts/gpu_round_check.py [--budget SECONDS]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
T0 = time.time()
BUDGET = float(sys.argv[sys.argv.index("--budget") + 1]) if "--budget" in sys.argv else 660.0
summary = {"stages": []}


def left():
    return BUDGET - (time.time() - T0)


def run(name, cmd, timeout, env=None, log=None):
    timeout = min(timeout, max(5.0, left()))
    e = dict(os.environ)
    e.update(env or {})
    t = time.time()
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=timeout)
        rc, out = r.returncode, r.stdout + "\n--- stderr ---\n" + r.stderr[-6000:]
    except subprocess.TimeoutExpired as ex:
        rc, out = 124, ((ex.stdout or b"").decode(errors="replace") if isinstance(ex.stdout, bytes) else (ex.stdout or "")) + "\nTIMEOUT"
    open(os.path.join(OUT, (log or name) + ".log"), "w").write(out)
    summary["stages"].append({"name": name, "rc": rc, "seconds": round(time.time() - t, 1), "env": env or {}})
    json.dump(summary, open(os.path.join(OUT, "round_check.json"), "w"), indent=1)
    print("[%6.1fs] %-28s rc=%d (%.1fs)" % (time.time() - T0, name, rc, time.time() - t), flush=True)
    return rc, out


def json_line(out):
    for l in reversed(out.splitlines()):
        if l.startswith("{") and '"metric"' in l:
            try:
                return json.loads(l)
            except Exception:  # noqa: BLE001
                pass
    return None


py = sys.executable
# 1) parity of the new pieces (persistent correlation kernel vs the one-tile kernel and the oracle; concurrent lanes)
rc_new, _ = run("tests_new", [py, "-m", "pytest", "tests/test_gpu_tc.py", "tests/test_gpu_pair.py", "-q", "-x", "-s",
                              "-k", "persistent or concurrent"], 240)
rc_pair, _ = run("tests_corr_neigh_pair", [py, "-m", "pytest", "tests/test_gpu_ops.py", "-q", "-x", "-k", "corr_neigh"], 120)
pair_ok = rc_pair == 0
summary["corr_neigh_pair_ok"] = pair_ok
# 2) everything that touches the correlation / the pair path, with the new pieces as the defaults
new_env = {"RF_CORR_V2": "1", "RF_CORR_NEIGH_PAIR": "1" if pair_ok else "0"}
rc_v2, _ = run("tests_new_defaults", [py, "-m", "pytest", "tests/test_gpu_matching.py", "tests/test_gpu_tc.py", "tests/test_gpu_pair.py",
                                      "-q", "-x", "-k", "not conv2d"], 300, env=new_env)
v2_ok = rc_new == 0 and rc_v2 == 0
summary["v2_parity_ok"] = v2_ok
# 3) the correlation call alone, both sequences
rc, out = run("kernel_bench_corr2", [py, "scripts/kernel_bench.py", "corr2", "10"], 120)
summary["kernel_bench_corr2"] = [l for l in out.splitlines() if l.startswith("corr_mutual_nn") or l.startswith("identical")]
# 4) bench matrix
results = {}
old_env = {"RF_CORR_V2": "0", "RF_CORR_NEIGH_PAIR": "0", "RF_LANES": "1"}
cur = {"RF_CORR_V2": "1" if v2_ok else "0", "RF_CORR_NEIGH_PAIR": "1" if (pair_ok and v2_ok) else "0"}
matrix = [("old lanes=1", old_env)]
if cur != {k: old_env[k] for k in cur}:
    matrix += [("new lanes=1", dict(cur, RF_LANES="1"))]
matrix += [("%s lanes=%d" % ("new" if v2_ok else "old", n), dict(cur, RF_LANES=str(n))) for n in (2, 3, 4)]
for name, env in matrix:
    if left() < 200:
        break
    rc, out = run("bench " + name, [py, "bench.py", "--no-cpu-baseline"], 150, env=env, log="bench_" + name.replace(" ", "_").replace("=", ""))
    line = json_line(out)
    if rc == 0 and line:
        results[name] = {"env": env, "value": line["value"], "e2e": line["e2e"]["value"], "ms_per_step": line["ms_per_step"],
                         "corr_ms": line["roofline"]["ms_per_launch"], "frac": line["roofline"]["frac"], "clocks": line["clocks"]}
summary["bench_matrix"] = results
json.dump(summary, open(os.path.join(OUT, "round_check.json"), "w"), indent=1)
best = max(results.items(), key=lambda kv: kv[1]["value"]) if results else None
best_env = best[1]["env"] if best else old_env
summary["best"] = best[0] if best else None
# 5) the final line with the CPU baseline, in the best configuration
rc, out = run("bench_final", [py, "bench.py"], 200, env=best_env)
line = json_line(out)
if line:
    open(os.path.join(OUT, "bench_line_final.json"), "w").write(json.dumps(line) + "\n")
# 6) ncu --set full of the correlation kernel (one launch)
if v2_ok and left() > 60:
    run("ncu_corr_pipe", ["ncu", "--set", "full", "--clock-control", "none", "--import-source", "on", "--kernel-name-base", "mangled",
                          "-k", "regex:tc_corr_pipe", "-c", "1", "-f", "-o", os.path.join(OUT, "prof_corr_pipe"),
                          py, "scripts/kernel_bench.py", "corr2v", "1"], 150)
# 7) per-kernel shares of a step (CUPTI), default engine
if left() > 45:
    run("profile_step_f16", [py, "scripts/profile_step.py", "f16"], 120, env=best_env)
summary["total_seconds"] = round(time.time() - T0, 1)
json.dump(summary, open(os.path.join(OUT, "round_check.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))
