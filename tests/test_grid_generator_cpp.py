"""The C++ ball generator of the reference-facing mirror (include/dasm/grid_generator.h) against the Python one
(dealii-asm_b200/grid.py): same vertices, cells, support points and parent maps (host only, compiled with g++)."""
import importlib
import os
import subprocess

import numpy as np
import pytest

from __graft_entry__ import ROOT

grid = importlib.import_module("dealii-asm_b200.grid")

SRC = r"""
#include <cstdio>
#include "%s/include/dasm/grid_generator.h"
int main(int argc, char **argv)
{
  const unsigned int L = (unsigned int)std::atoi(argv[1]);
  const auto m = dasm::GridGenerator::hyper_ball(L);
  const auto p = dasm::GridGenerator::ball_parents(L);
  FILE *f = std::fopen(argv[2], "wb");
  const long long sizes[4] = {m.n_vertices(), m.n_cells(), (long long)m.support.size(), (long long)p.size()};
  std::fwrite(sizes, sizeof(long long), 4, f);
  std::fwrite(m.vertices.data(), sizeof(double), m.vertices.size(), f);
  std::fwrite(m.cells.data(), sizeof(std::uint32_t), m.cells.size(), f);
  std::fwrite(m.support.data(), sizeof(double), m.support.size(), f);
  std::fwrite(p.data(), sizeof(std::uint32_t), p.size(), f);
  std::fclose(f);
  return 0;
}
"""


@pytest.fixture(scope="module")
def generator(tmp_path_factory):
    d = tmp_path_factory.mktemp("gridgen")
    src = d / "gen.cc"
    src.write_text(SRC % ROOT)
    exe = d / "gen"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), str(src)])
    return exe, d


@pytest.mark.parametrize("L", [0, 1, 3])
def test_cpp_ball_equals_python_ball(generator, L):
    exe, d = generator
    out = d / ("ball%d.bin" % L)
    subprocess.check_call([str(exe), str(L), str(out)])
    raw = open(out, "rb").read()
    nv, nc, ns, npar = np.frombuffer(raw, dtype=np.int64, count=4)
    off = 32
    vertices = np.frombuffer(raw, dtype=np.float64, count=3 * nv, offset=off).reshape(nv, 3)
    off += 24 * nv
    cells = np.frombuffer(raw, dtype=np.uint32, count=8 * nc, offset=off).reshape(nc, 8)
    off += 32 * nc
    support = np.frombuffer(raw, dtype=np.float64, count=ns, offset=off).reshape(nc, 27, 3)
    off += 8 * ns
    parents = np.frombuffer(raw, dtype=np.uint32, count=npar, offset=off)
    g = grid.hyper_ball(L)
    assert nc == 32 * 8 ** L and nv == len(g["vertices"])
    assert np.array_equal(cells, g["cells"])
    assert np.allclose(vertices, g["vertices"], rtol=0, atol=1e-14)
    assert np.allclose(support, g["support"], rtol=0, atol=1e-14)
    if L > 0:
        assert np.array_equal(parents, grid.ball_parents(L))
