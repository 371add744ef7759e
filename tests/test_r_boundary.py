"""The R side of the boundary (r/ccgp_shim.c, r/ccgp.R), checked without R (there is none in the image):
the shim goes through a real compiler against stand-in R headers (tests/r_stub/) and the real include/ccgp.h,
and the R file is lexed: balanced delimiters, every .Call target registered with the right argument count,
every reference closure of SURVEY 8b defined."""
import os
import re
import subprocess

from r_lex import check_balanced, strip_strings_and_comments

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "r", "ccgp_shim.c")
RFILE = os.path.join(ROOT, "r", "ccgp.R")


def registered_routines():
    src = open(SHIM).read()
    return {m.group(1): int(m.group(2)) for m in re.finditer(r"CALLDEF\((ccgp_R_\w+),\s*(\d+)\)", src)}


def test_shim_compiles_against_the_c_abi_header():
    cmd = ["gcc", "-fsyntax-only", "-std=c99", "-Wall", "-Wextra", "-Wno-cast-function-type", "-Werror",
           "-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include"), SHIM]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr


def test_every_shim_routine_is_registered_with_its_arity():
    src = open(SHIM).read()
    defined = {}
    for m in re.finditer(r"^SEXP (ccgp_R_\w+)\(([^)]*)\)\s*\{", src, re.M):
        defined[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
    reg = registered_routines()
    assert defined == reg, (set(defined) ^ set(reg), {k: (defined.get(k), reg.get(k)) for k in defined if defined.get(k) != reg.get(k)})
    assert "R_registerRoutines" in src and "R_init_ccgp_shim" in src


def _call_sites(code):
    """(.Call target, number of arguments after the name) for every .Call( in lexed R code."""
    out = []
    for m in re.finditer(r"\.Call\(", code):
        i = m.end()
        depth, args, cur = 1, [], []
        while depth:
            ch = code[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                args.append("".join(cur)); cur = []
            else:
                cur.append(ch)
            i += 1
        args.append("".join(cur))
        out.append(len(args) - 1)
    return out


def test_r_wrappers_are_consistent_with_the_shim():
    src = open(RFILE).read()
    assert check_balanced(src) is None, check_balanced(src)
    reg = registered_routines()
    names = re.findall(r'\.Call\("(ccgp_R_\w+)"', src)
    counts = _call_sites(strip_strings_and_comments(src))
    assert len(names) == len(counts) and len(names) >= 15
    for name, n in zip(names, counts):
        assert name in reg, name
        assert reg[name] == n, (name, reg[name], n)
    assert set(reg) - set(names) == set(), "registered but never called from ccgp.R: %s" % (set(reg) - set(names))


def test_reference_closures_are_defined_and_drivers_left_alone():
    code = strip_strings_and_comments(open(RFILE).read())
    defined = set(re.findall(r"^([A-Za-z.][\w.]*)\s*<-\s*function", code, re.M))
    for name in ["logpost", "logpost.batch", "Mixed.corr.matrix", "Mixed.corr.vec", "cross.corr.matrix", "likeli.hyperpars",
                 "choose.hyperpars", "predict.post.batch", "prediction.table", "Entropy", "Augmented.Mixed.Entropy",
                 "Batch.Entropy.optim", "Entropy.optim", "entropy.batch", "entropy.argmin", "Metro.multichain",
                 "kmedoids.design", "subset.logdet.batch", "loglik.argmax", "ccgp.init", "ccgp.use.script"]:
        assert name in defined, name
    # the reference's own drivers keep running on top of the shadowed closures: they must NOT be redefined
    for name in ["predict.post", "prediction", "factors", "factors.frame", "Metro", "compare.GP", "Combined.GP.fit", "beta.MLE"]:
        assert name not in defined, name
    # one prior line per script, as cited
    for alias in ["A", "I", "M", "G", "V", "H", "D1", "D2"]:
        assert re.search(r"\b%s\s*=\s*list\(family" % alias, code), alias
