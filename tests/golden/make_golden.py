"""Transcribes golden vectors out of /root/reference into small committed fixtures."""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def evs(path):
    out = []
    with open(path) as f:
        for line in f:
            m = re.match(r"\s+- (min|max) ev: (\S+)", line)
            if m:
                out.append((m.group(1), float(m.group(2))))
    return [{"min": out[i][1], "max": out[i + 1][1]} for i in range(0, len(out), 2)]


def main():
    small = os.path.join(REF, "tests/element_centered_preconitioners/small")
    g = {}
    for name in ["dummy_mg_chebyshev_fdm_1_post", "dummy_mg_chebyshev_fdm_1_pre", "dummy_mg_chebyshev_fdm_1_symm",
                 "dummy_mg_chebyshev_fdm_1_none", "dummy_mg_chebyshev_fdm_3", "dummy_chebyshev_diagonal",
                 "dummy_mg_chebyshev_asm"]:
        g[name] = evs(os.path.join(small, name + ".output"))
    json.dump(g, open(os.path.join(HERE, "chebyshev_fdm_estimates.json"), "w"), indent=1)
    lines = open(os.path.join(REF, "indices_overlap_01.output")).read().split("\n")
    open(os.path.join(HERE, "indices_overlap_01.output.txt"), "w").write("\n".join(lines[:320]) + "\n")
    rows = []
    for line in open(os.path.join(REF, "subdivided_hyper_cube_balanced_01.output")):
        p = line.split()
        if len(p) == 6:
            rows.append([int(p[0]), int(p[1]), int(p[2]), int(p[3]), int(p[4]), float(p[5])])
    json.dump(rows, open(os.path.join(HERE, "subdivided_hyper_cube_balanced_01.json"), "w"))
    # reduced_access_01.result / reduced_access_02.result: command lines and the printed local vectors as integer lists
    for name in ("reduced_access_01", "reduced_access_02"):
        cases = []
        for block in open(os.path.join(REF, name + ".result")).read().split("./" + name)[1:]:
            lines = block.strip().splitlines()
            cases.append({"args": [int(a) for a in lines[0].split()], "local": [int(x) for l in lines[1:] for x in l.split()]})
        json.dump(cases, open(os.path.join(HERE, name + ".json"), "w"))


if __name__ == "__main__":
    main()
