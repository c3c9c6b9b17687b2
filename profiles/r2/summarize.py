#!/usr/bin/env python
"""Markdown tables from the bench lines committed in this directory.  usage: python profiles/r2/summarize.py"""
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    p = os.path.join(HERE, name)
    if not os.path.exists(p):
        return None
    try:
        txt = [l for l in open(p).read().strip().splitlines() if l.startswith("{")]
        return json.loads(txt[-1])
    except Exception:
        return None


def row(name, d):
    c = d["config"]
    sh = d.get("time_share", {})
    pr = c.get("per_round_ms", [])
    kn = c.get("per_round_knn_ms", [])
    extra = ""
    if d.get("n_gpus", 1) > 1:
        extra = " bit-identical to 1 GPU: %s" % d.get("bit_identical_to_1gpu")
    return "| `%s` | %d | %.1f | %.3f | %.1f | %s | %s | %s | %s | %.0f |%s" % (
        name, d.get("n_gpus", 1), d["value"], d["ms_per_step"], d["e2e"]["value"],
        ("%.2f" % pr[0]) if pr else "-", ("%.2f" % pr[-1]) if pr else "-",
        ("%.2f / %.2f" % (kn[0], kn[-1])) if kn else "-",
        " / ".join("%.0f %%" % (100 * sh.get(k, 0)) for k in ("knn", "select", "lm_eval", "lm_other")),
        c.get("setup_ms_excluded", 0), extra)


def main():
    print("| file | GPUs | iter/s | ms/step | e2e iter/s | round 0 ms | last round ms | knn ms (round 0 / last) | share knn / select / lm_eval / lm_other | setup ms |")
    print("|---|---:|---:|---:|---:|---:|---:|---|---|---:|")
    for p in sorted(glob.glob(os.path.join(HERE, "*.json"))):
        d = load(os.path.basename(p))
        if not d or d.get("impl") == "reference" or "config" not in d or "per_round_ms" not in d["config"]:
            continue
        print(row(os.path.basename(p), d))
    print()
    for p in sorted(glob.glob(os.path.join(HERE, "*.json"))):
        d = load(os.path.basename(p))
        if d and d.get("impl") == "reference":
            cb = d["cpu_baseline"]
            st = cb.get("single_thread", {})
            print("reference arm `%s`: %.4f iter/s on %d threads (corr %.2f s + LM %.2f s per round), single thread %.4f iter/s; rounds %d; inliers %s; LM iterations %s"
                  % (os.path.basename(p), d["value"], cb["cores"], cb["corr_s_per_round"], cb["lm_s_per_round"], st.get("value", float("nan")), d["steps"],
                     d["inliers_per_round"][:3], d["lm_iterations_per_round"][:6]))


if __name__ == "__main__":
    main()
