"""Test infrastructure (build container only): a stand-in for the third-party ``xir`` package, built on
``strawberryfields_b200.io`` -- the XIR counterpart of ``tests/blackbird_facade.py``.

The reference converts programs through ``xir.Program`` / ``xir.Statement`` / ``xir.Declaration`` and reads
scripts with ``xir.parse_script`` (``strawberryfields/io/xir_io.py:33-330``, ``io/__init__.py:145-237``).  Here
the containers are plain classes with the attributes the reference touches; ``parse_script`` and
``Program.serialize`` go through our XIR parser / writer (gate definitions are already expanded by the parser,
so ``Program.gates`` is empty).  Time-domain programs (``_type_: tdm``) are parsed (looped-over arrays kept by name)
for the reference's converter; running or writing them is out of scope.
"""
import sys
import types
from decimal import Decimal

import numpy as np

from strawberryfields_b200 import io as bio


class DecimalComplex(complex):
    """``xir.DecimalComplex(real, imag)``: only its value matters to the reference (``xir_io._listr``)"""

    def __new__(cls, real="0", imag="0"):
        return super().__new__(cls, float(Decimal(real)), float(Decimal(imag)))


class Statement:
    def __init__(self, name, params, wires):
        self.name, self.params, self.wires = name, params, tuple(wires)


class Declaration:
    def __init__(self, name, type_, params=None, wires=None):
        self.name, self.type_, self.params, self.wires = name, type_, list(params or []), tuple(wires or ())


class Program:
    def __init__(self, version="0.1.0"):
        self.version = version
        self.options, self.constants = {}, {}
        self.statements = []
        self.declarations = {"gate": [], "out": [], "func": [], "obs": []}
        self.gates = {}      # user-defined gates: expanded at parse time by strawberryfields_b200.io

    def add_option(self, key, value):
        self.options[key] = value

    def add_constant(self, key, value):
        self.constants[key] = value

    def add_statement(self, stmt):
        self.statements.append(stmt)

    def add_declaration(self, decl):
        self.declarations[decl.type_].append(decl)

    @property
    def wires(self):
        return {w for s in self.statements for w in s.wires}

    def search(self, decl_type, attr, name):
        raise KeyError(name)   # no unexpanded gate definitions survive the parser

    def serialize(self):
        if self.options.get("_type_") == "tdm":
            raise NotImplementedError("time-domain XIR programs are out of scope for b200fock's io")
        opts = {k: v for k, v in self.options.items() if k != "_name_"}
        prog = bio.CircuitProgram(name=self.options.get("_name_"), version=self.version, options=opts)
        for s in self.statements:
            params = s.params
            prog.operations.append({"op": s.name, "modes": list(s.wires),
                                    "args": [] if isinstance(params, dict) else [_param(a) for a in params],
                                    "kwargs": {k: _param(a) for k, a in params.items()} if isinstance(params, dict) else {}})
        text = bio.dumps(prog, ir="xir")
        decls = ["%s %s%s[%s];" % (d.type_, d.name, "(%s)" % ", ".join(d.params) if d.params else "",
                                   ", ".join(map(str, d.wires)))
                 for kind in ("gate", "out") for d in self.declarations[kind]]
        return "\n".join(decls + ([""] if decls and text else []) + ([text] if text else []))


def _param(a):
    """what ``io.to_xir`` puts into a statement -> our parameters: strings are free-parameter expressions"""
    if isinstance(a, str) and a not in ("real", "complex", "square"):
        try:
            return bio._eval(a, {bio._XIR_NAMES: True})
        except (bio.ProgramSyntaxError, NameError):
            return a
    if isinstance(a, list):
        return np.array(a)
    return a


def _plain(v):
    """our values -> what xir hands out: names / expressions as strings, arrays as nested lists"""
    if isinstance(v, bio.Parameter):
        return bio._fmt_xir(v)
    if isinstance(v, np.ndarray):
        return v.tolist()
    return v


def parse_script(script, eval_pi=False, use_floats=True, **kwargs):
    prog = bio._loads_xir(script)
    out = Program()
    if prog.name is not None:
        out.add_option("_name_", prog.name)
    for k, v in prog.options.items():
        out.add_option(k, v)
    for k, v in prog.variables.items():     # time-domain programs: the looped-over arrays (xir_io.py:160-170)
        out.add_constant(k, v)
    for op in prog.operations:
        params = {k: _plain(v) for k, v in op["kwargs"].items()} if op["kwargs"] else [_plain(a) for a in op["args"]]
        out.add_statement(Statement(op["op"], params, tuple(op["modes"])))
    return out


def install():
    mod = sys.modules.get("xir")
    if mod is None:
        mod = sys.modules["xir"] = types.ModuleType("xir")
    for name, obj in (("Program", Program), ("Statement", Statement), ("Declaration", Declaration),
                      ("DecimalComplex", DecimalComplex), ("parse_script", parse_script)):
        setattr(mod, name, obj)
    mod.__version__ = "b200fock-io-facade"
    return mod
