#!/usr/bin/env python3
"""Generate rnaelem_b200/csrc/energy_data.inc from ViennaRNA-2.0 format parameter text.

The reference embeds two third-party ViennaRNA parameter sets (Turner 2004, Andronescu 2007) as C string
literals (RNAelem/energy_param.hpp:652-660 includes rna_turner2004.par / rna_andronescu2007.par).  This tool
reads those data files where they lie (it is run in the build container, where /root/reference exists),
applies the *reading rules* of the reference's parser (energy_param.hpp:159-183 get_array, 423-502 the scalar
sections, 519-640 the section table: which sub-block of each table is read, which words are skipped) and
writes the resulting integer tables (dcal/mol, sentinel for INF) as compact C arrays in our own layout:

    only the sub-blocks the reference actually reads, flattened in reading order.

The conversion to log-Boltzmann weights (-E*10/kT, Vienna 'smooth' on multi/exterior/dangles) is done at load
time by energy_tables.cpp (product) and relem_oracle.c (oracle) -- not here.

Usage: python tools/gen_energy_data.py [/root/reference/RNAelem] > rnaelem_b200/csrc/energy_data.inc
"""
import sys, os, re

INF = 1000000  # sentinel for "INF" (zeroL)


def load_lines(path):
    """Undo the C-string-literal wrapping: each physical line is  "....\\n"  ."""
    out = []
    for raw in open(path):
        raw = raw.rstrip("\n")
        m = re.match(r'^\s*"(.*)"\s*$', raw)
        if not m:
            if raw.strip() == "":
                continue
            raise SystemExit("unexpected line in %s: %r" % (path, raw))
        s = m.group(1).replace("\\t", "\t")
        assert s.endswith("\\n"), raw
        out.append(s[:-2])
    return out


class Stream:
    def __init__(self, lines):
        self.lines, self.pos = lines, 0

    def getline(self):
        if self.pos >= len(self.lines):
            return None
        s = self.lines[self.pos]
        self.pos += 1
        return s


def atoi(w):
    m = re.match(r"\s*([+-]?\d+)", w)
    return int(m.group(1)) if m else 0


def get_array(st, size):
    """energy_param.hpp:159-183: fill up to `size` ints; a line shorter than 2 chars ends the fill,
    a word containing '/*' ends the line."""
    vals = []
    while len(vals) < size:
        s = st.getline()
        if s is None or len(s) < 2:
            break
        for w in s.split():
            if len(vals) >= size:
                break
            if "/*" in w:
                break
            if w == "INF":
                vals.append(INF)
            elif w == "DEF":
                vals.append(-50)
            else:
                vals.append(atoi(w))
    vals += [INF] * (size - len(vals))
    return vals


def param_type(s):
    if len(s) == 0 or s[0] != "#":
        return None
    w = s.split()
    return w[1] if len(w) > 1 else None


def parse(lines):
    T = {}
    st = Stream(lines)
    # pass 1 (read_only_misc, energy_param.hpp:504-517): lxc37 if the Misc line has > 4 words
    T["lxc37"] = 107.856
    while True:
        s = st.getline()
        if s is None:
            break
        if param_type(s) == "Misc":
            while True:
                s = st.getline()
                if s is None or s == "":
                    break
                if "*" in s:
                    continue
                w = s.split()
                if len(w) > 4:
                    T["lxc37"] = float(w[4])
            break
    st = Stream(lines)
    while True:
        s = st.getline()
        if s is None:
            break
        t = param_type(s)
        if t == "stack":  # [7][7] rows 1..6, cols 1..6
            T["stack"] = [get_array(st, 6) for _ in range(6)]
        elif t in ("mismatch_hairpin", "mismatch_interior", "mismatch_interior_1n", "mismatch_interior_23"):
            T[t] = [get_array(st, 25) for _ in range(6)]  # types 1..6, [5][5]
        elif t in ("mismatch_multi", "mismatch_exterior"):
            T[t] = [get_array(st, 25) for _ in range(7)]  # types 1..7 (dim 8 read), smoothed
        elif t in ("dangle5", "dangle3"):
            T[t] = [get_array(st, 5) for _ in range(7)]  # types 1..7, smoothed
        elif t == "int11":
            T[t] = [[get_array(st, 25) for _ in range(7)] for _ in range(7)]
        elif t == "int21":
            T[t] = [[get_array(st, 125) for _ in range(7)] for _ in range(7)]
        elif t == "int22":  # types 1..6 x 1..6, bases 1..4 each, 4 numbers per get_array call
            T[t] = [[[[[get_array(st, 4) for _ in range(4)] for _ in range(4)] for _ in range(4)]
                     for _ in range(6)] for _ in range(6)]
        elif t in ("hairpin", "bulge", "interior"):
            T[t] = get_array(st, 31)
        elif t == "NINIO":
            while True:
                s = st.getline()
                if s is None or s == "":
                    break
                if "*" in s:
                    continue
                w = s.split()
                T["ninio"] = [atoi(w[0]), atoi(w[2])]
                break
        elif t == "ML_params":
            while True:
                s = st.getline()
                if s is None or s == "":
                    break
                if "*" in s:
                    continue
                w = s.split()
                T["ml"] = [atoi(w[0]), atoi(w[2]), atoi(w[4])]  # base, closing, intern
                break
        elif t == "Misc":
            while True:
                s = st.getline()
                if s is None or s == "":
                    break
                if "*" in s:
                    continue
                w = s.split()
                T["term_au"] = atoi(w[2])
        elif t in ("Triloops", "Tetraloops", "Hexaloops"):
            items = []
            while True:
                s = st.getline()
                if s is None or s == "":
                    break
                if "*" in s:
                    continue
                w = s.split()
                items.append((w[0], atoi(w[1])))
            T[t] = items
    return T


def flat(x):
    if isinstance(x, list):
        r = []
        for e in x:
            r += flat(e)
        return r
    return [x]


def emit(name, T, out):
    def arr(key, cname):
        v = flat(T[key])
        out.append("static const int %s_%s[%d] = {" % (name, cname, len(v)))
        for i in range(0, len(v), 16):
            out.append("  " + ",".join(str(a) for a in v[i:i + 16]) + ",")
        out.append("};")

    arr("stack", "stack")
    arr("mismatch_hairpin", "mm_h")
    arr("mismatch_interior", "mm_i")
    arr("mismatch_interior_1n", "mm_1ni")
    arr("mismatch_interior_23", "mm_23i")
    arr("mismatch_multi", "mm_m")
    arr("mismatch_exterior", "mm_ext")
    arr("dangle5", "dangle5")
    arr("dangle3", "dangle3")
    arr("int11", "int11")
    arr("int21", "int21")
    arr("int22", "int22")
    arr("hairpin", "hairpin")
    arr("bulge", "bulge")
    arr("interior", "interior")
    out.append("static const int %s_scalars[6] = {%d,%d,%d,%d,%d,%d}; /* ninio f, ninio max, ml base, ml closing, ml intern, termAU */"
               % (name, T["ninio"][0], T["ninio"][1], T["ml"][0], T["ml"][1], T["ml"][2], T["term_au"]))
    out.append("static const double %s_lxc37 = %r;" % (name, T["lxc37"]))
    for key, cname in (("Triloops", "tri"), ("Tetraloops", "tetra"), ("Hexaloops", "hexa")):
        items = T.get(key, [])
        out.append("static const int %s_n%s = %d;" % (name, cname, len(items)))
        out.append("static const char* const %s_%s_seq[%d] = {%s};" %
                   (name, cname, max(1, len(items)), ",".join('"%s"' % a for a, _ in items) or '""'))
        out.append("static const int %s_%s_e[%d] = {%s};" %
                   (name, cname, max(1, len(items)), ",".join(str(b) for _, b in items) or "0"))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/RNAelem"
    out = ["/* GENERATED by tools/gen_energy_data.py -- do not edit.",
           " * ViennaRNA-2.0 format parameter sets (Turner 2004; Andronescu 2007), integer dcal/mol,",
           " * %d = INF.  Only the sub-blocks the reference parser reads, flattened in reading order:" % INF,
           " *   stack[6][6] mm_*[6|7][5][5] dangle*[7][5] int11[7][7][5][5] int21[7][7][5][5][5]",
           " *   int22[6][6][4][4][4][4] hairpin/bulge/interior[31]. */",
           "#define RELEM_EINF %d" % INF]
    for name, fn in (("T2004", "rna_turner2004.par"), ("A2007", "rna_andronescu2007.par")):
        T = parse(load_lines(os.path.join(ref, fn)))
        emit("RELEM_" + name, T, out)
    print("\n".join(out))


if __name__ == "__main__":
    main()
