#!/usr/bin/env python3
"""Extract the literal golden arrays held by the reference's own unit tests.

Run in the development container only (it reads /root/reference, which does not exist on the
GPU box).  The reference package itself cannot be imported here (xarray / rasterio / json_checker
are absent), so instead of running its tests we parse them with ``ast`` and evaluate every
assignment (and every ``pytest.mark.parametrize`` table) whose right-hand side is a pure numpy
literal.  The result is committed as ``reference_goldens.npz`` (+ ``reference_goldens.json`` with
the file:line provenance of every key) and is what ``tests/test_oracle_goldens.py`` pins the
oracle against.

Key format:  ``<relative test file>::<function or Class.method>::<variable>[#n]``
(``#n`` = n-th re-assignment of the same name inside the function, 0-based, omitted for the first;
``@k`` = k-th numpy literal nested inside a right-hand side that is not itself a pure literal, e.g. the arrays
inside an ``xr.Dataset(...)`` call, in ``ast.walk`` order),
and for parametrised tests ``<file>::<function>[<case index>]::<argname>``.
"""
from __future__ import annotations

import ast
import json
import os
import sys

import numpy as np

REF_TESTS = os.environ.get("PANDORA_REFERENCE", "/root/reference") + "/tests"
HERE = os.path.dirname(os.path.abspath(__file__))

# (file, function qualname) pairs that pin the hot path
TARGETS = [
    ("common.py", "matching_cost_tests_setup"),
    ("test_matching_cost/test_matching_cost_census.py", "test_census_cost"),
    ("test_matching_cost/test_matching_cost_census.py", "test_census"),
    ("test_matching_cost/test_matching_cost_census.py", "test_cmax"),
    ("test_matching_cost/test_matching_cost_sad.py", "*"),
    ("test_matching_cost/test_matching_cost_ssd.py", "*"),
    ("test_matching_cost/test_matching_cost_zncc.py", "*"),
    ("test_matching_cost/test_matching_cost.py", "*"),
    ("test_aggregation.py", "*"),
    ("test_disparity.py", "*"),
    ("test_filter.py", "*"),
    ("test_criteria.py", "*"),
    # SURVEY.md 8(f) "next" rows: refinement, cross-checking, confidence
    ("test_refinement.py", "*"),
    ("test_validation.py", "*"),
    ("test_confidence/conftest.py", "*"),
    ("test_confidence/test_ambiguity.py", "*"),
    ("test_confidence/test_risk.py", "*"),
]


class _Param:
    def __init__(self, *values, id=None, marks=None):  # noqa: A002
        self.values = values
        self.id = id


class _FakePytest:
    param = _Param

    class mark:  # noqa: N801
        @staticmethod
        def parametrize(*a, **k):
            return ("parametrize", a, k)


def _storable(val):
    if isinstance(val, np.ndarray):
        return val.dtype.kind in "fiub"
    if isinstance(val, (int, float, np.integer, np.floating)):
        return True
    if isinstance(val, (list, tuple)) and len(val) > 0:
        try:
            arr = np.asarray(val)
        except Exception:  # ragged
            return False
        return arr.dtype.kind in "fiub"
    return False


def _eval(node, env):
    code = compile(ast.Expression(body=node), "<golden>", "eval")
    return eval(code, {"__builtins__": {"abs": abs, "float": float, "int": int, "range": range, "len": len}}, env)


def _target_name(tgt):
    if isinstance(tgt, ast.Name):
        return tgt.id
    if isinstance(tgt, ast.Attribute) and isinstance(tgt.value, ast.Name) and tgt.value.id == "self":
        return "self." + tgt.attr
    return None


def _walk_function(fn, qual, relfile, out, prov, module_env):
    env = dict(module_env)
    seen = {}
    for node in ast.walk(fn):
        pass  # (ordering comes from the statement walk below)

    def visit(stmts):
        for st in stmts:
            if isinstance(st, ast.Assign) and len(st.targets) == 1:
                name = _target_name(st.targets[0])
                if name is None:
                    continue
                try:
                    val = _eval(st.value, env)
                except Exception:
                    # not a pure literal (e.g. ``xr.Dataset({... np.array([...]) ...})``): keep the numpy literals
                    # nested inside it, in source order, as ``<name>@<k>``
                    k = 0
                    for sub in ast.walk(st.value):
                        if isinstance(sub, ast.Call) and ast.unparse(sub.func) in ("np.array", "np.asarray", "np.full", "np.arange"):
                            try:
                                sval = _eval(sub, env)
                            except Exception:
                                continue
                            if _storable(sval):
                                n = seen.get(name, 0)
                                key = f"{relfile}::{qual}::{name}" + (f"#{n}" if n else "") + f"@{k}"
                                out[key] = np.asarray(sval)
                                prov[key] = f"tests/{relfile}:{sub.lineno}"
                                k += 1
                    if k:
                        seen[name] = seen.get(name, 0) + 1
                    continue
                env[name.replace("self.", "self_")] = val
                if isinstance(st.targets[0], ast.Name):
                    env[name] = val
                if _storable(val):
                    n = seen.get(name, 0)
                    seen[name] = n + 1
                    key = f"{relfile}::{qual}::{name}" + (f"#{n}" if n else "")
                    out[key] = np.asarray(val)
                    prov[key] = f"tests/{relfile}:{st.lineno}"
            elif isinstance(st, (ast.With, ast.For, ast.If, ast.Try)):
                visit(getattr(st, "body", []))
                visit(getattr(st, "orelse", []))

    visit(fn.body)


def _parametrize(fn, qual, relfile, out, prov, module_env):
    for dec in fn.decorator_list:
        if not (isinstance(dec, ast.Call) and ast.unparse(dec.func).endswith("parametrize")):
            continue
        try:
            names = _eval(dec.args[0], module_env)
            table = _eval(dec.args[1], module_env)
        except Exception:
            continue
        if isinstance(names, str):
            names = [s.strip() for s in names.split(",")]
        for i, case in enumerate(table):
            values = case.values if isinstance(case, _Param) else case
            if len(names) == 1 and not isinstance(values, (tuple, list)):
                values = (values,)
            if isinstance(case, _Param) and case.id is not None:
                out_id = f"{relfile}::{qual}[{i}]::__id__"
                out[out_id] = np.frombuffer(case.id.encode(), dtype=np.uint8)
                prov[out_id] = f"tests/{relfile}:{dec.lineno}"
            for nm, val in zip(names, values):
                if _storable(val):
                    key = f"{relfile}::{qual}[{i}]::{nm}"
                    out[key] = np.asarray(val)
                    prov[key] = f"tests/{relfile}:{dec.lineno}"
                elif isinstance(val, dict):
                    # a parameter dictionary (e.g. the masks of TestCvMasked): ``<argname>.<key>[.<subkey>]``
                    def store(prefix, dic):
                        for dk, dv in dic.items():
                            if isinstance(dv, dict):
                                store(f"{prefix}.{dk}", dv)
                            elif _storable(dv):
                                key = f"{relfile}::{qual}[{i}]::{prefix}.{dk}"
                                out[key] = np.asarray(dv)
                                prov[key] = f"tests/{relfile}:{dec.lineno}"
                            elif isinstance(dv, str):
                                key = f"{relfile}::{qual}[{i}]::{prefix}.{dk}"
                                out[key] = np.frombuffer(dv.encode(), dtype=np.uint8)
                                prov[key] = f"tests/{relfile}:{dec.lineno}"

                    store(nm, val)


def _reference_constants():
    """Module-level integer constants of src/pandora/constants.py (validity-mask bits), exposed to the
    evaluated test literals as ``cst``."""
    path = os.path.join(os.path.dirname(REF_TESTS), "src", "pandora", "constants.py")
    ns = {}
    for st in ast.parse(open(path).read()).body:
        if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
            try:
                ns[st.targets[0].id] = _eval(st.value, ns)
            except Exception:
                pass
    return type("cst", (), ns)


def main():
    out, prov = {}, {}
    cst = _reference_constants()
    for relfile, want in TARGETS:
        path = os.path.join(REF_TESTS, relfile)
        tree = ast.parse(open(path).read())
        module_env = {"np": np, "pytest": _FakePytest, "n": np.nan, "cst": cst, "Affine": lambda *a, **k: None}
        # module-level simple constants (e.g. ``n = np.nan``)
        for st in tree.body:
            if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                try:
                    module_env[st.targets[0].id] = _eval(st.value, module_env)
                except Exception:
                    pass

        def handle(fn, qual):
            if want != "*" and want != qual and want != qual.split(".")[-1]:
                return
            _parametrize(fn, qual, relfile, out, prov, module_env)
            _walk_function(fn, qual, relfile, out, prov, module_env)

        for st in tree.body:
            if isinstance(st, ast.FunctionDef):
                handle(st, st.name)
            elif isinstance(st, ast.ClassDef):
                for sub in st.body:
                    if isinstance(sub, ast.FunctionDef):
                        handle(sub, f"{st.name}.{sub.name}")
    for name, val in vars(cst).items():
        if isinstance(val, int):
            out[f"constants.py::{name}"] = np.asarray(val)
            prov[f"constants.py::{name}"] = "src/pandora/constants.py"
    np.savez_compressed(os.path.join(HERE, "reference_goldens.npz"), **out)
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as fh:
        json.dump(prov, fh, indent=0, sort_keys=True)
    print(f"{len(out)} golden arrays written")
    return 0


if __name__ == "__main__":
    sys.exit(main())
