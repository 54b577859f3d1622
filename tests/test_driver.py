"""The C++ driver with the reference's CLI / JSON / `>>` output (matrix_free_loop_08.likwid.cc:390-395)."""
import json
import os
import subprocess

import pytest

from __graft_entry__ import ROOT

DRIVER = os.path.join(ROOT, "drivers", "matrix_free_loop_08")


def test_driver_binary_and_headers_exist():
    for f in ("include/dasm.h", "include/dasm/operator.h", "include/dasm/preconditioners.h", "include/dasm/precondition.h",
              "include/dasm/json.h", "drivers/matrix_free_loop_08.cc"):
        assert os.path.exists(os.path.join(ROOT, f)), f


@pytest.mark.gpu
def test_matrix_free_loop_08_output(tmp_path):
    if not os.path.exists(DRIVER):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "drivers")])
    labels = "vmult post-1-c symm-1-c add-1-g-p-n symm-2-g-p-n cheby-3-2-symm-1-c cheby-2-3-diag"
    cfg = {"dim": 3, "number type": "double", "fe degree": 3, "n subdivisions": 14, "preconditioner types": labels,
           "n repetitions": 2, "use cartesian mesh": True}
    p = tmp_path / "in.json"
    p.write_text(json.dumps(cfg))
    out = subprocess.run([DRIVER, str(p)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.splitlines() if l.startswith(">>")]
    assert [l[1] for l in lines] == labels.split()
    n_cells = (8, 4, 4)  # s = 14 -> n_refine 2, subdivisions (2,1,1)
    n_dofs = 1
    for c in n_cells:
        n_dofs *= c * 3
    for l in lines:
        assert int(l[2]) == n_dofs          # n_dofs
        assert float(l[4]) > 0              # time
        assert l[5] == "8" and l[6] == "3"  # sizeof(Number), degree
    assert [int(l[3]) for l in lines] == [2, 2, 2, 2, 2, 6, 4]  # repetitions * chebyshev degree
    # unknown mapping type -> error exit like the reference's AssertThrow
    cfg["mapping type"] = "bogus"
    p.write_text(json.dumps(cfg))
    out = subprocess.run([DRIVER, str(p)], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "is not known" in out.stderr
