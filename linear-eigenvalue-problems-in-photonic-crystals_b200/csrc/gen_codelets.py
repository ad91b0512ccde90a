#!/usr/bin/env python3
"""Generate pcb_codelets.cuh: straight-line in-register complex DFT codelets.

    python gen_codelets.py            # writes pcb_codelets.cuh next to this file
    python gen_codelets.py --selftest # checks every codelet against numpy.fft

A codelet ``dftR_f`` / ``dftR_b`` transforms R complex128 values held in registers
(``cplx v[R]``, natural order in, natural order out), forward = exp(-2 pi i jk/R),
backward = exp(+2 pi i jk/R), both unnormalised.  Radices 2,3,4,5 are explicit
butterflies; composite radices are built by Good-Thomas (coprime factors, no twiddles)
or Cooley-Tukey with literal twiddle constants (trivial ones folded away).  All index
maps are resolved at generation time, so the emitted code is pure arithmetic on named
scalars which nvcc keeps in registers and contracts into FMAs.
"""
import math
import os
import sys

RADICES = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16]
FACTOR = {6: (2, 3), 8: (2, 4), 9: (3, 3), 10: (2, 5), 12: (3, 4), 15: (3, 5), 16: (4, 4)}


class Gen:
    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self, expr):
        name = f"t{self.n}"
        self.n += 1
        self.lines.append(f"const double {name} = {expr};")
        return name

    # complex helpers on (re, im) name pairs -------------------------------------------
    def add(self, a, b):
        return (self.tmp(f"{a[0]} + {b[0]}"), self.tmp(f"{a[1]} + {b[1]}"))

    def sub(self, a, b):
        return (self.tmp(f"{a[0]} - {b[0]}"), self.tmp(f"{a[1]} - {b[1]}"))

    def mul_i(self, a, s):
        """a * (s*i), s = +1 or -1, without emitting code (sign folded into names)."""
        return (neg(a[1]), a[0]) if s > 0 else (a[1], neg(a[0]))

    def cmul(self, a, w):
        wr, wi = w
        eps = 1e-15
        if abs(wi) < eps and abs(wr - 1) < eps:
            return a
        if abs(wi) < eps and abs(wr + 1) < eps:
            return (neg(a[0]), neg(a[1]))
        if abs(wr) < eps and abs(wi - 1) < eps:
            return self.mul_i(a, +1)
        if abs(wr) < eps and abs(wi + 1) < eps:
            return self.mul_i(a, -1)
        if abs(abs(wr) - abs(wi)) < eps:   # 45 degrees: c*(+-1 +- i)
            c = lit(abs(wr))
            sr, si = (1 if wr > 0 else -1), (1 if wi > 0 else -1)
            # (ar + i ai) * c (sr + i si) = c[(sr ar - si ai) + i (si ar + sr ai)]
            re = self.tmp(f"{c} * ({sgn(sr)}{a[0]} {'-' if si > 0 else '+'} {a[1]})")
            im = self.tmp(f"{c} * ({sgn(si)}{a[0]} {'+' if sr > 0 else '-'} {a[1]})")
            return (re, im)
        re = self.tmp(f"{lit(wr)} * {a[0]} - {lit(wi)} * {a[1]}")
        im = self.tmp(f"{lit(wr)} * {a[1]} + {lit(wi)} * {a[0]}")
        return (re, im)

    # butterflies -----------------------------------------------------------------------
    def dft(self, xs, s):
        """s = -1 forward, +1 backward."""
        n = len(xs)
        if n == 1:
            return list(xs)
        if n == 2:
            return [self.add(xs[0], xs[1]), self.sub(xs[0], xs[1])]
        if n == 3:
            a, b, c = xs
            t = self.add(b, c)
            d = self.sub(b, c)
            x0 = self.add(a, t)
            m = (self.tmp(f"{a[0]} - 0.5 * {t[0]}"), self.tmp(f"{a[1]} - 0.5 * {t[1]}"))
            h = lit(math.sqrt(3.0) / 2.0)
            # X1 = m + s*i*h*d ; X2 = m - s*i*h*d
            hd = (self.tmp(f"{h} * {d[0]}"), self.tmp(f"{h} * {d[1]}"))
            ihd = self.mul_i(hd, s)
            return [x0, self.add(m, ihd), self.sub(m, ihd)]
        if n == 4:
            a, b, c, d = xs
            t0, t1 = self.add(a, c), self.sub(a, c)
            t2, t3 = self.add(b, d), self.sub(b, d)
            it3 = self.mul_i(t3, s)
            return [self.add(t0, t2), self.add(t1, it3), self.sub(t0, t2), self.sub(t1, it3)]
        if n == 5:
            a, b, c, d, e = xs
            t1, t2 = self.add(b, e), self.add(c, d)
            t3, t4 = self.sub(b, e), self.sub(c, d)
            c1, c2 = lit(math.cos(2 * math.pi / 5)), lit(math.cos(4 * math.pi / 5))
            s1, s2 = lit(math.sin(2 * math.pi / 5)), lit(math.sin(4 * math.pi / 5))
            x0 = (self.tmp(f"{a[0]} + ({t1[0]} + {t2[0]})"), self.tmp(f"{a[1]} + ({t1[1]} + {t2[1]})"))
            r1 = tuple(self.tmp(f"{a[k]} + {c1} * {t1[k]} + {c2} * {t2[k]}") for k in (0, 1))
            r2 = tuple(self.tmp(f"{a[k]} + {c2} * {t1[k]} + {c1} * {t2[k]}") for k in (0, 1))
            i1 = tuple(self.tmp(f"{s1} * {t3[k]} + {s2} * {t4[k]}") for k in (0, 1))
            i2 = tuple(self.tmp(f"{s2} * {t3[k]} - {s1} * {t4[k]}") for k in (0, 1))
            j1, j2 = self.mul_i(i1, s), self.mul_i(i2, s)
            return [x0, self.add(r1, j1), self.add(r2, j2), self.sub(r2, j2), self.sub(r1, j1)]
        n1, n2 = FACTOR[n]
        out = [None] * n
        if math.gcd(n1, n2) == 1:      # Good-Thomas
            u = pow(n2, -1, n1)
            v = pow(n1, -1, n2)
            A = [[xs[(n2 * a + n1 * b) % n] for b in range(n2)] for a in range(n1)]
            C = [self.dft([A[a][b] for a in range(n1)], s) for b in range(n2)]     # C[b][k1]
            for k1 in range(n1):
                row = self.dft([C[b][k1] for b in range(n2)], s)
                for k2 in range(n2):
                    out[(n2 * u * k1 + n1 * v * k2) % n] = row[k2]
        else:                           # Cooley-Tukey
            C = [self.dft([xs[a * n2 + b] for a in range(n1)], s) for b in range(n2)]   # C[b][k1]
            for k1 in range(n1):
                tw = []
                for b in range(n2):
                    ang = s * 2 * math.pi * (k1 * b) / n
                    tw.append(self.cmul(C[b][k1], (math.cos(ang), math.sin(ang))))
                row = self.dft(tw, s)
                for k2 in range(n2):
                    out[k1 + n1 * k2] = row[k2]
        return out


def lit(x):
    return repr(float(x))


def sgn(s):
    return "" if s > 0 else "-"


def neg(name):
    return name[1:] if name.startswith("-") else "-" + name


def fix(expr):
    """Clean up '+ -x' / '- -x' produced by folded negations."""
    return (expr.replace("+ -", "- ").replace("- -", "+ ").replace("(-", "(-").replace("* -t", "* (-1.0) * t"))


def codelet(R, s):
    g = Gen()
    xs = [(f"v[{i}].x", f"v[{i}].y") for i in range(R)]
    # read inputs into scalars first so in-place output is safe
    ins = []
    for i in range(R):
        ins.append((g.tmp(xs[i][0]), g.tmp(xs[i][1])))
    outs = g.dft(ins, s)
    body = [fix(l) for l in g.lines]
    for i, (re, im) in enumerate(outs):
        body.append(fix(f"v[{i}].x = {re}; v[{i}].y = {im};"))
    return body


def count_ops(body):
    return sum(l.count(" + ") + l.count(" - ") + l.count(" * ") for l in body)


def emit():
    out = ["// GENERATED by gen_codelets.py -- do not edit.", "#pragma once", "",
           "// In-register complex DFT codelets (natural order in/out, unnormalised).", ""]
    for R in RADICES:
        for s, tag in ((-1, "f"), (+1, "b")):
            body = codelet(R, s)
            out.append(f"// radix {R} {'forward' if s < 0 else 'backward'}: ~{count_ops(body)} real ops")
            out.append(f"PCB_HD void dft{R}_{tag}(cplx* __restrict__ v) {{")
            out += ["  " + l for l in body]
            out.append("}")
            out.append("")
    out.append("template <int R, int DIR> struct Dft;   // DIR = -1 forward, +1 backward")
    out.append("template <int DIR> struct Dft<1, DIR> { PCB_HD static void run(cplx*) {} };")
    for R in RADICES:
        out.append(f"template <> struct Dft<{R}, -1> {{ PCB_HD static void run(cplx* v) {{ dft{R}_f(v); }} }};")
        out.append(f"template <> struct Dft<{R}, +1> {{ PCB_HD static void run(cplx* v) {{ dft{R}_b(v); }} }};")
    out.append("")
    return "\n".join(out)


def selftest():
    import numpy as np
    rng = np.random.default_rng(0)
    for R in RADICES:
        for s in (-1, 1):
            body = codelet(R, s)
            x = rng.standard_normal(R) + 1j * rng.standard_normal(R)

            class C:
                pass
            v = []
            for z in x:
                c = C()
                c.x, c.y = z.real, z.imag
                v.append(c)
            env = {"v": v}
            for l in body:
                l = l.replace("const double ", "")
                for stmt in l.split(";"):
                    if stmt.strip():
                        exec(stmt.strip(), {}, env)
            got = np.array([c.x + 1j * c.y for c in v])
            want = np.fft.fft(x) if s < 0 else np.fft.ifft(x) * R
            err = np.max(np.abs(got - want)) / np.max(np.abs(want))
            assert err < 1e-14, (R, s, err)
            print(f"radix {R:2d} dir {s:+d}: ops {count_ops(body):4d}  err {err:.1e}")


if __name__ == "__main__":
    if "--selftest" in sys.argv:
        selftest()
    else:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pcb_codelets.cuh")
        with open(path, "w") as f:
            f.write(emit())
        print("wrote", path)
