#!/usr/bin/env python
"""Extracts the reference's OWN Monte Carlo results for the S = 1/2 cases on the accelerated path from the
golden outputs of its regression tests (loop.op, extras/*/*.op: ALPS `loop --evaluate` text, one block per
task) into tests/golden/ref_runs.json, so that the restatement in oracle/ (and through it the GPU engine) is
pinned to numbers the reference itself printed -- including the ALGORITHM-specific "Number of Clusters",
which no exact diagonalisation can supply.  Statistical goldens: value and error as printed (1024 or 4096
sweeps; errors flagged "NOT CONVERGED" by ALPS are kept with that flag).
Run in the build container only (reads /root/reference); the JSON travels."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
FILES = ["loop.op", "extras/transmag/transmag.op", "extras/gap/gap.op", "extras/corrlen/corrlen.op",
         "extras/top/top.op", "extras/localsus/localsus.op"]
KEEP = ["Energy", "Energy Density", "Number of Clusters", "Magnetization^2", "Magnetization^4",
        "Staggered Magnetization^2", "Staggered Magnetization^4", "Susceptibility", "Staggered Susceptibility",
        "Generalized Magnetization^2", "Generalized Susceptibility", "Stiffness", "Transverse Magnetization",
        "Energy^2", "Specific Heat", "Binder Ratio of Magnetization", "Binder Ratio of Staggered Magnetization"]
RES = re.compile(r"^([A-Za-z|][^:]*): (-?[0-9.eE+-]+|nan|inf) \+/- ([0-9.eE+-]+|inf|nan)(.*)$")
PAR = re.compile(r'^([A-Za-z_][A-Za-z0-9_\[\] ]*) = (.*);$')


def blocks(path):
    """[(first line number, params, results)] -- a block = parameter lines, then result lines"""
    out, params, results, start, in_results = [], {}, {}, 1, False
    for n, ln in enumerate(open(path), 1):
        ln = ln.rstrip("\n")
        m = PAR.match(ln)
        if m and not ln.startswith(" "):
            if in_results:                      # a new block begins
                out.append((start, params, results))
                params, results, in_results, start = {}, {}, False, n
            params[m.group(1).strip()] = m.group(2).strip().strip('"')
            continue
        m = RES.match(ln)
        if m:
            in_results = True
            name = m.group(1)
            if name in KEEP:
                results[name] = dict(line=n, value=float(m.group(2)), error=float(m.group(3)),
                                     converged="NOT CONVERGED" not in m.group(4))
    out.append((start, params, results))
    return out


def num(s):
    if "/" in s:
        a, b = s.split("/")
        return float(a) / float(b)
    return float(s)


def main():
    runs = []
    for f in FILES:
        for start, p, r in blocks(os.path.join(REF, f)):
            algo = p.get("ALGORITHM", "")
            if algo not in ("loop; path integral", "loop; sse", "diagonalization") or not r:
                continue
            if p.get("local_S", "1/2") not in ("1/2", "0.5") or "h" in p or "D" in p and num(p["D"]) != 0:
                continue                        # S > 1/2, longitudinal field, single-ion term: not on the path
            lat = p.get("LATTICE", "")
            if lat not in ("chain lattice", "site", "simple cubic lattice", "alternating chain lattice"):
                continue
            L = int(num(p["L"])) if "L" in p and "/" not in p.get("T", "") else (int(num(p["L"])) if "L" in p else 1)
            T = p["T"]
            T = 1.0 / L if T.replace(" ", "") == "1/L" else num(T)
            J = num(p.get("J", "0"))
            if lat == "alternating chain lattice":   # extras/transmag: Jz0 = Jz1, Jxy0 = Jxy1, Gamma0 = -Gamma1
                jz, jxy, gamma = num(p["Jz0"]), num(p["Jxy0"]), abs(num(p["Gamma0"]))
                assert num(p["Jz1"]) == jz and num(p["Jxy1"]) == jxy and num(p["Gamma1"]) == -num(p["Gamma0"])
                lat = "chain lattice"               # the staggered field is uniform after the sublattice rotation
            else:
                jz, jxy, gamma = num(p.get("Jz", str(J))), num(p.get("Jxy", str(J))), num(p.get("Gamma", "0"))
            if lat == "site":
                L, jz, jxy = 1, 0.0, 0.0
            runs.append(dict(source=f"{f}:{start}", algorithm=algo, lattice=lat, L=L, T=T, Jz=jz, Jxy=jxy, Gamma=gamma,
                             improved="DISABLE_IMPROVED_ESTIMATOR" not in p, sweeps=int(p.get("SWEEPS", "0")),
                             extra=sorted(k for k in p if k.startswith("MEASURE")), results=r))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_runs.json")
    json.dump(runs, open(path, "w"), indent=1)
    for x in runs:
        print(x["source"], x["algorithm"], x["lattice"], x["L"], x["T"], x["Jz"], x["Jxy"], x["Gamma"],
              "improved" if x["improved"] else "normal", sorted(x["results"]))


if __name__ == "__main__":
    main()
