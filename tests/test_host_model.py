"""Host-side model construction of the product (rnaelem_b200/csrc/host_model.cpp) against the reference's parse and
automaton dumps.  The functions are reached through the C ABI introspection calls; on the GPU-less box the ABI is
provided by the host-emulation build of the same sources (tests/emu), on the GPU box test_gpu_parity covers the
real library."""
import json
import os

import numpy as np
import pytest

import caselib
import rnaelem_b200 as rb
from test_oracle import TABLES, _valid_mask, _hmm_check


@pytest.mark.parametrize("which", ["T2004", "A2007"])
def test_energy_tables(which, emu_lib):
    g = np.load(os.path.join(caselib.GOLDEN, "tables_%s.npz" % which))
    ctx = rb.Context(0, lib=emu_lib)
    ctx.set_energy("~%s~" % which, 50, 30, 1e-4, 0)
    for name in TABLES:
        a, b = ctx.energy_get(name), g[name]
        n = min(len(a), len(b))
        mask = _valid_mask(name, n)
        np.testing.assert_array_equal(a[:n][mask], b[:n][mask], err_msg=name)
    for name, key in (("triloop", "triloops"), ("tetraloop", "tetraloops"), ("hexaloop", "hexaloops")):
        k = len(str(g[key]).split())
        np.testing.assert_array_equal(ctx.energy_get(name)[:k], g[name][:k])


def test_automata(emu_lib):
    g = json.load(open(os.path.join(caselib.GOLDEN, "hmm.json")))
    ctx = rb.Context(0, lib=emu_lib)
    for pat, d in g.items():
        ctx.set_pattern(pat)
        assert (ctx.M, ctx.S) == (d["M"], d["S"]), pat
        _hmm_check(ctx.hmm_get, d)
        assert ctx.row_sizes == d["theta_rows"]


def test_bad_patterns_are_refused(emu_lib):
    ctx = rb.Context(0, lib=emu_lib)
    for pat in ["", "(.", ".)", "(x)"]:
        with pytest.raises(rb.RelemError):
            ctx.set_pattern(pat)


def test_param_file_reader_equals_builtin(emu_lib, tmp_path):
    """A ViennaRNA-2.0 format text written from the built-in Turner2004 integers must parse back to the same tables
    (exercises parse_param_text with the reference's reading rules: INF/DEF words, comment words, block order)."""
    g = np.load(os.path.join(caselib.GOLDEN, "tables_T2004.npz"))
    ctx = rb.Context(0, lib=emu_lib)
    ctx.set_energy("~T2004~", 50, 30, 1e-4, 0)
    kT = (37 + 273.15) * 1.98717

    def ints(name, smooth=False):
        v = ctx.energy_get(name)
        out = []
        for x in v:
            out.append("INF" if x == -np.inf else str(int(round(-x * kT / 10.))))
        return out
    lines = ["## RNAfold parameter file v2.0", ""]
    st = np.array(ints("stack")).reshape(7, 7)
    lines += ["# stack"] + [" ".join(st[a, 1:]) + "   /* row */" for a in range(1, 7)] + [""]
    for sec, name in (("mismatch_hairpin", "mismatch_h"), ("mismatch_interior", "mismatch_i"),
                      ("mismatch_interior_1n", "mismatch_1ni"), ("mismatch_interior_23", "mismatch_23i")):
        t = np.array(ints(name)).reshape(7, 5, 5)
        lines += ["# " + sec] + [" ".join(t[a, b]) for a in range(1, 7) for b in range(5)] + [""]
    for sec, name in (("hairpin", "hairpin"), ("bulge", "bulge"), ("interior", "internal")):
        v = ints(name)
        lines += ["# " + sec] + [" ".join(v[k:k + 10]) for k in range(0, 31, 10)] + [""]
    lines += ["# NINIO", "/* Ninio = MIN(max, m*|n1-n2| */", "/*\t\t    m\t  m_dH     max  max_dH\t*/", "\t       60    320   300     0", ""]
    lines += ["# Misc", "/* all parameters are pairs of 'energy enthalpy' */", "/*    DuplexInit     TerminalAU      LXC */",
              "   410  360    50  370 107.856000    0", ""]
    lines += ["# Triloops", "CAACG   680  2370", "GUUAC   690  1080", "", "#END"]
    p = tmp_path / "mini.par"
    p.write_text("\n".join(lines) + "\n")
    c2 = rb.Context(0, lib=emu_lib)
    c2.set_energy(str(p), 50, 30, 1e-4, 0)
    for name in ("stack", "mismatch_h", "mismatch_i", "mismatch_1ni", "mismatch_23i", "hairpin", "bulge", "internal",
                 "ninio", "term_au"):
        a, b = c2.energy_get(name), ctx.energy_get(name)
        if name in ("mismatch_1ni",):
            a, b = a[25:], b[25:]
        np.testing.assert_allclose(a, b, rtol=0, atol=0, err_msg=name)
    np.testing.assert_array_equal(c2.energy_get("triloop"), ctx.energy_get("triloop"))
    assert c2.energy_get("lxc37")[0] == 107.856
