"""Program I/O for ``b200fock``: Blackbird (``.xbb``) and XIR (``.xir``) scripts in and out,
and an on-disk checkpoint of the simulated state -- without the ``blackbird`` / ``xir`` packages.

The reference loads stored programs with ``sf.load`` / ``sf.loads`` and writes them with ``sf.save``
(``/root/reference/strawberryfields/io/__init__.py:67,145,169``), going through the third-party parsers
``blackbird`` and ``xir`` (``io/blackbird_io.py:30,164``, ``io/xir_io.py:64,218``).  Neither is in this
image, and the GPU box has no Strawberry Fields either, so this module carries its own small parsers for
the parts of both languages a Fock-backend program uses:

* Blackbird: ``name`` / ``version`` / ``target dev (opt = val)`` header, scalar and array variable
  declarations (``float x = 0.3``, ``complex array U[4, 4] =`` + indented rows), comments, and operation
  lines ``Op(args, key=val) | modes``, and ``for <type> <var> in start:stop[:step]`` / ``in [..]`` loops over
  an indented block of operations (unrolled; the variable may appear in arguments, array subscripts and
  mode lists).  Template parameters (``{par}``) and measured parameters (``q0`` = the last outcome of mode 0)
  stay symbolic (:class:`Parameter`) until ``run(backend, args={..})`` binds / feeds them forward -- what
  ``par_convert`` + ``par_evaluate`` do in the reference (``parameters.py:162-275``, ``engine.py:427-444``).
  ``type`` / ``include`` statements (time-domain programs, out of scope) are refused with
  ``NotImplementedError``.
* XIR: ``options: .. end;`` and ``constants: .. end;`` blocks, ``use`` / declaration statements (ignored),
  statements ``Op(args, key: val) | [wires];`` and gate definitions ``gate Name(params)[wires]: .. end;``,
  expanded in place where they are used (``xir.get_expanded_statements``, ``xir_io.py:89``).

A loaded script is a :class:`CircuitProgram`: the list of operations in the form the reference's
converters consume (``{"op", "args", "kwargs", "modes"}``, ``blackbird_io.py:46-75``).  ``calls()`` lowers
it to ``BaseFock`` backend calls -- what ``LocalEngine._run_program`` makes after the ``fock`` compiler has
decomposed the program (``engine.py:422-457``): primitive gates map one to one; ``Xgate``, ``Zgate``,
``Pgate``, ``CXgate``, ``CZgate``, ``Fouriergate``, ``sMZgate`` use the reference's decompositions
(``ops.py:1726-1732,2149-2158,2209-2216``) and ``Interferometer(U)`` the rectangular (Clements) mesh in the
reference's gate order (``ops.py:2655-2718``).  ``run(backend)`` executes it and returns the measurement
samples; ``to_sf()`` builds a ``strawberryfields.Program`` when Strawberry Fields is importable and ``from_sf(prog)``
goes the other way (``io.to_blackbird``, ``blackbird_io.py:164-232``).

State checkpoints (``save_state`` / ``load_state``) are ``.npz`` files with the ket or density matrix in
the reference's layout (``backend.py:50-56``) plus cutoff, purity and mode count; a state object of a
sharded circuit gathers its shards first (checkpoints of states that do not fit one host are written per
rank by ``ShardedCircuit.save_shard`` / ``load_shard``).
"""
from __future__ import annotations

import ast
import math
import os
import re

import numpy as np

HBAR = 2.0  # strawberryfields.hbar default; Xgate / Zgate displacements are x / sqrt(2 hbar)

_FUNCS = {
    "sqrt": np.sqrt, "sin": np.sin, "cos": np.cos, "tan": np.tan, "exp": np.exp, "log": np.log,
    "arcsin": np.arcsin, "arccos": np.arccos, "arctan": np.arctan, "arctan2": np.arctan2,
    "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan, "sinh": np.sinh, "cosh": np.cosh,
    "tanh": np.tanh, "arcsinh": np.arcsinh, "arccosh": np.arccosh, "arctanh": np.arctanh, "abs": abs,
}
_CONSTS = {"pi": math.pi, "PI": math.pi, "e": math.e, "True": True, "False": False, "true": True, "false": False,
           "None": None, "none": None}


class ProgramSyntaxError(ValueError):
    """A script that is not valid Blackbird / XIR (``blackbird.BlackbirdSyntaxError`` in the reference)."""


class Parameter:
    """A symbolic gate argument: a free parameter (Blackbird ``{name}``), a measured one (``q3`` = the latest
    outcome of mode 3) or arithmetic on them.  The reference keeps these as sympy expressions over
    ``FreeParameter`` / ``MeasuredParameter`` atoms (``parameters.py:318-420``); here it is a small expression
    tree -- ``("free", name)``, ``("measured", mode)``, ``("bin", symbol, a, b)``, ``("neg", a)``,
    ``("call", function, args)`` with numbers as leaves -- that :meth:`evaluate` walks once every atom has a
    value, :meth:`substitute` partially binds and ``text`` writes back as Blackbird."""

    __array_ufunc__ = None    # numpy scalars on the left hand side defer to __rmul__ & co.
    _OPS = {"+": lambda x, y: x + y, "-": lambda x, y: x - y, "*": lambda x, y: x * y, "/": lambda x, y: x / y,
            "**": lambda x, y: x ** y}

    def __init__(self, node):
        self.node = node

    @classmethod
    def free(cls, name):
        return cls(("free", name))

    @classmethod
    def measured(cls, mode):
        return cls(("measured", int(mode)))

    @classmethod
    def call(cls, name, args):
        return cls(("call", name, tuple(a.node if isinstance(a, Parameter) else a for a in args)))

    # ---- tree walks
    @classmethod
    def _atoms(cls, node, out):
        if isinstance(node, tuple):
            if node[0] in ("free", "measured"):
                out.add(node)
            else:
                for child in (node[2] if node[0] == "call" else node[1:]):
                    cls._atoms(child, out)
        return out

    @property
    def atoms(self):
        return self._atoms(self.node, set())

    @property
    def free_names(self):
        return sorted(k for kind, k in self.atoms if kind == "free")

    @property
    def measured_modes(self):
        return sorted(k for kind, k in self.atoms if kind == "measured")

    @classmethod
    def _value(cls, node, look, funcs):
        if not isinstance(node, tuple):
            return node
        if node[0] in ("free", "measured"):
            return look(*node)
        if node[0] == "bin":
            return cls._OPS[node[1]](cls._value(node[2], look, funcs), cls._value(node[3], look, funcs))
        if node[0] == "neg":
            return -cls._value(node[1], look, funcs)
        return funcs[node[1]](*[cls._value(a, look, funcs) for a in node[2]])

    def evaluate(self, lookup, funcs=None):
        """``lookup(kind, key)`` supplies the atoms (numbers, or sympy symbols for ``to_sf``)."""
        return self._value(self.node, lookup, _FUNCS if funcs is None else funcs)

    @classmethod
    def _subst(cls, node, values):
        if not isinstance(node, tuple):
            return node
        if node[0] == "free":
            return values.get(node[1], node)
        if node[0] == "measured":
            return node
        if node[0] == "call":
            args = tuple(cls._subst(a, values) for a in node[2])
            return node[:2] + (args,) if any(isinstance(a, tuple) for a in args) else _FUNCS[node[1]](*args)
        kids = [cls._subst(c, values) for c in (node[2:] if node[0] == "bin" else node[1:])]
        if any(isinstance(c, tuple) for c in kids):
            return (node[:2] if node[0] == "bin" else node[:1]) + tuple(kids)
        return cls._OPS[node[1]](*kids) if node[0] == "bin" else -kids[0]

    def substitute(self, values):
        """bind some free parameters; a number if nothing symbolic is left, else a new :class:`Parameter`"""
        node = self._subst(self.node, values)
        return Parameter(node) if isinstance(node, tuple) else node

    @classmethod
    def _text(cls, node, top=True):
        if not isinstance(node, tuple):
            t = _fmt(node)
            return "(%s)" % t if not top and (t.startswith("-") or isinstance(node, (complex, np.complexfloating))) else t
        if node[0] == "free":
            return "{%s}" % node[1]
        if node[0] == "measured":
            return "q%d" % node[1]
        if node[0] == "call":
            return "%s(%s)" % (node[1], ", ".join(cls._text(a) for a in node[2]))
        t = "-%s" % cls._text(node[1], False) if node[0] == "neg" else \
            "%s %s %s" % (cls._text(node[2], False), node[1], cls._text(node[3], False))
        return t if top else "(%s)" % t

    @property
    def text(self):
        return self._text(self.node)

    # ---- arithmetic
    def _bin(self, sym, other, swap=False):
        o = other.node if isinstance(other, Parameter) else other
        return Parameter(("bin", sym, o, self.node) if swap else ("bin", sym, self.node, o))

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __pow__(self, o): return self._bin("**", o)
    def __rpow__(self, o): return self._bin("**", o, True)
    def __neg__(self): return Parameter(("neg", self.node))
    def __pos__(self): return self

    def __repr__(self):
        return "Parameter(%s)" % self.text


_XIR_NAMES = "__xir_names__"     # env key: unknown names are free parameters (XIR scripts)
_FREE_PREFIX = "_b200_free_"   # ``{name}`` is rewritten to this prefix + name so that ``ast`` can parse the expression


def _symbolic(v):
    if isinstance(v, Parameter):
        return True
    if isinstance(v, (list, tuple)):
        return any(_symbolic(x) for x in v)
    return isinstance(v, np.ndarray) and v.dtype == object and any(_symbolic(x) for x in v.flat)


def _resolve(v, lookup, funcs=None):
    if isinstance(v, Parameter):
        return v.evaluate(lookup, funcs)
    if isinstance(v, (list, tuple)):
        return type(v)(_resolve(x, lookup, funcs) for x in v)
    if isinstance(v, np.ndarray) and v.dtype == object:
        out = np.array([_resolve(x, lookup, funcs) for x in v.flat], dtype=object).reshape(v.shape)
        try:
            return out.astype(complex) if any(isinstance(x, complex) for x in out.flat) else out.astype(float)
        except (TypeError, ValueError):
            return out
    return v


def _eval(expr, env):
    """Evaluate a Blackbird / XIR parameter expression: numbers (``1+2j`` included), declared variables,
    ``pi``, arithmetic and the usual real functions.  Anything else is refused (no ``eval``)."""
    expr = expr.strip()
    if not expr:
        raise ProgramSyntaxError("empty expression")
    expr = re.sub(r"\{\s*([A-Za-z_]\w*)\s*\}", lambda m: _FREE_PREFIX + m.group(1), expr)
    try:
        tree = ast.parse(expr, mode="eval")
    except SyntaxError as exc:
        raise ProgramSyntaxError("cannot parse expression %r" % expr) from exc

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant):
            if isinstance(node.value, (int, float, complex, bool, str)) or node.value is None:
                return node.value
        elif isinstance(node, ast.Name):
            if node.id in env:
                return env[node.id]
            if node.id in _CONSTS:
                return _CONSTS[node.id]
            if node.id.startswith(_FREE_PREFIX):
                return Parameter.free(node.id[len(_FREE_PREFIX):])
            if env.get(_XIR_NAMES):   # XIR has no {name} syntax: a name that is not a constant is a free parameter
                return Parameter.free(node.id)      # (io.to_xir writes them so, xir_io.py:287-303)
            if re.fullmatch(r"q\d+", node.id):
                return Parameter.measured(int(node.id[1:]))
            raise NameError("name %r is not defined in the script" % node.id)
        elif isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            v = ev(node.operand)
            return -v if isinstance(node.op, ast.USub) else v
        elif isinstance(node, ast.BinOp):
            a, b = ev(node.left), ev(node.right)
            if isinstance(node.op, ast.Add):
                return a + b
            if isinstance(node.op, ast.Sub):
                return a - b
            if isinstance(node.op, ast.Mult):
                return a * b
            if isinstance(node.op, ast.Div):
                return a / b
            if isinstance(node.op, ast.Pow):
                return a ** b
        elif isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id in _FUNCS \
                and not node.keywords:
            vals = [ev(a) for a in node.args]
            if any(isinstance(v, Parameter) for v in vals):
                return Parameter.call(node.func.id, vals)
            return _FUNCS[node.func.id](*vals)
        elif isinstance(node, (ast.List, ast.Tuple)):
            return [ev(e) for e in node.elts]
        elif isinstance(node, ast.Subscript) and isinstance(node.value, ast.Name) and node.value.id in env:
            idx = ev(node.slice)
            return np.asarray(env[node.value.id])[tuple(idx) if isinstance(idx, list) else idx]
        raise ProgramSyntaxError("unsupported expression %r" % expr)

    return ev(tree)


def _split_top(text, sep=","):
    """split at ``sep`` outside brackets / parentheses"""
    out, depth, cur = [], 0, []
    for ch in text:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        out.append("".join(cur))
    return out


def _parse_args(text, env, kw_sep):
    args, kwargs = [], {}
    for item in _split_top(text):
        item = item.strip()
        m = re.match(r"^([A-Za-z_]\w*)\s*%s(?!=)\s*(.+)$" % re.escape(kw_sep), item, re.S)
        if m:
            kwargs[m.group(1)] = _eval(m.group(2), env)
        else:
            if kwargs:
                raise ProgramSyntaxError("positional argument after a keyword argument: %r" % item)
            args.append(_eval(item, env))
    return args, kwargs


def _parse_modes(text, env=None):
    text = text.strip()
    if text and text[0] in "[(":
        if text[-1] not in "])":
            raise ProgramSyntaxError("unbalanced mode list %r" % text)
        text = text[1:-1]
    modes = []
    for tok in _split_top(text):
        tok = tok.strip()
        if not tok:
            continue
        if re.fullmatch(r"\d+", tok):
            modes.append(int(tok))
            continue
        if not env:
            raise ProgramSyntaxError("mode index %r is not an integer" % tok)
        val = _eval(tok, env)  # loop variables and integer arithmetic on them (``| [m, m + 1]``)
        if isinstance(val, bool) or not isinstance(val, (int, np.integer)) or val < 0:
            raise ProgramSyntaxError("mode index %r is not a non-negative integer" % tok)
        modes.append(int(val))
    if not modes:
        raise ProgramSyntaxError("an operation needs at least one mode")
    return modes


_OP_LINE = re.compile(r"^([A-Za-z_]\w*)\s*(?:\((.*)\))?\s*\|\s*(.+)$", re.S)


class CircuitProgram:
    """A loaded circuit: header fields and the operation list (``{"op", "args", "kwargs", "modes"}``)."""

    def __init__(self, name=None, version="1.0", target=None, operations=None, options=None):
        self.name = name
        self.version = version
        self.target = target or {"name": None, "options": {}}
        self.operations = list(operations or [])
        self.options = dict(options or {})     # XIR options (cutoff_dim, shots, ..)
        self._bare = set()                     # indices of operations written without an argument list
        self.programtype = {"name": None, "options": {}}   # Blackbird ``type`` line / XIR ``_type_`` option
        self.variables = {}                    # time-domain programs: the looped-over arrays p0, p1, .. by name

    @property
    def num_subsystems(self):
        return 1 + max((m for op in self.operations for m in op["modes"]), default=-1)

    @property
    def run_options(self):
        opts = dict(self.target.get("options", {}))
        return {k: opts[k] for k in ("shots",) if k in opts} | {k: self.options[k] for k in ("shots",) if k in self.options}

    @property
    def backend_options(self):
        opts = dict(self.target.get("options", {}))   # blackbird_io.py:84-85 / xir_io.py:130-131
        return {k: opts[k] for k in ("cutoff_dim",) if k in opts} | {k: self.options[k] for k in ("cutoff_dim",) if k in self.options}

    # ------------------------------------------------------------------ symbolic parameters
    def _symbols(self):
        for op in self.operations:
            for v in list(op.get("args", [])) + list(op.get("kwargs", {}).values()):
                stack = [v]
                while stack:
                    x = stack.pop()
                    if isinstance(x, Parameter):
                        yield x
                    elif isinstance(x, (list, tuple)):
                        stack.extend(x)
                    elif isinstance(x, np.ndarray) and x.dtype == object:
                        stack.extend(x.flat)

    @property
    def free_parameters(self):
        """names of the template parameters (``{name}``) of the script, sorted"""
        return sorted({n for p in self._symbols() for n in p.free_names})

    @property
    def is_template(self):
        return bool(self.free_parameters)

    @property
    def has_feed_forward(self):
        return any(p.measured_modes for p in self._symbols())

    @staticmethod
    def _lookup(args, samples):
        def look(kind, key):
            if kind == "free":
                if args is None or key not in args:
                    raise ValueError("free parameter {%s} of the program has no value: pass args={%r: ..}" % (key, key))
                return args[key]
            if key not in samples:
                # the reference raises ParameterError from MeasuredParameter._eval_evalf (parameters.py:352-362)
                raise ValueError("parameter q%d is used before mode %d has been measured" % (key, key))
            return samples[key][-1]
        return look

    def bind(self, **values):
        """A copy with the given free parameters replaced by numbers (measured parameters stay symbolic)."""
        unknown = sorted(set(values) - set(self.free_parameters))
        if unknown:
            raise ValueError("the program has no free parameter(s) %s" % ", ".join(unknown))

        def sub(v):
            if isinstance(v, Parameter):
                return v.substitute(values)
            if isinstance(v, (list, tuple)):
                return type(v)(sub(x) for x in v)
            if isinstance(v, np.ndarray) and v.dtype == object:
                return _resolve(np.array([sub(x) for x in v.flat], dtype=object).reshape(v.shape), None)
            return v

        ops = [{"op": op["op"], "args": [sub(a) for a in op.get("args", [])],
                "kwargs": {k: sub(a) for k, a in op.get("kwargs", {}).items()}, "modes": list(op["modes"])}
               for op in self.operations]
        return CircuitProgram(self.name, self.version, self.target, ops, self.options)

    # ------------------------------------------------------------------ lowering to backend calls
    def _lower_op(self, op, look, cutoff=None):
        args = [_resolve(a, look) for a in op.get("args", [])]
        kwargs = {k: _resolve(a, look) for k, a in op.get("kwargs", {}).items()}
        return lower(op["op"], args, kwargs, op["modes"], cutoff)

    def calls(self, args=None, cutoff_dim=None):
        """``[(method, *args)]``: the ``BaseFock`` calls of the program (measurements included, in place).
        ``args`` binds template parameters; a program with measured parameters has no static call list
        (``run`` feeds the outcomes forward)."""
        if self.has_feed_forward:
            raise ValueError("the program uses measured parameters (q<mode>): its calls depend on the outcomes, use run()")
        look = self._lookup(args, {})
        D = cutoff_dim if cutoff_dim is not None else self.backend_options.get("cutoff_dim")
        out = []
        for op in self.operations:
            out.extend(self._lower_op(op, look, D))
        return out

    def run(self, backend, cutoff_dim=None, args=None, **begin_options):
        """``begin_circuit`` + every call; returns ``{mode: [outcomes]}`` like ``Result.samples_dict``.
        ``args = {name: value}`` binds the template parameters (``engine.run(prog, args=..)``,
        ``engine.py:304-306``); a measured parameter ``q<m>`` takes the latest outcome of mode m when the
        operation that uses it is reached (``engine.py:427-444`` + ``parameters.py:352-362``)."""
        D = cutoff_dim if cutoff_dim is not None else self.backend_options.get("cutoff_dim")
        if D is None:
            raise ValueError("Argument 'cutoff_dim' must be passed to the Fock backend")
        missing = sorted(set(self.free_parameters) - set(args or {}))
        if missing:
            raise ValueError("free parameter(s) %s of the program have no value: pass args={..}" % ", ".join(missing))
        backend.begin_circuit(self.num_subsystems, cutoff_dim=int(D), **begin_options)
        samples = {}
        look = self._lookup(args, samples)
        for op in self.operations:
            for call in self._lower_op(op, look, int(D)):
                ret = getattr(backend, call[0])(*call[1:-1], **call[-1]) if isinstance(call[-1], dict) else \
                    getattr(backend, call[0])(*call[1:])
                if call[0].startswith("measure_"):
                    modes = call[2] if call[0] == "measure_homodyne" else call[1]
                    modes = [modes] if isinstance(modes, int) else list(modes)
                    vals = np.asarray(ret).reshape(-1)
                    for m, v in zip(modes, vals):
                        samples.setdefault(m, []).append(v.item() if hasattr(v, "item") else v)
        return samples

    def to_sf(self):
        """The same program as a ``strawberryfields.Program`` (needs Strawberry Fields)."""
        import strawberryfields as sf
        from strawberryfields import ops

        prog = sf.Program(self.num_subsystems, name=self.name)
        funcs = {k: getattr(sf.math, k) for k in _FUNCS if hasattr(sf.math, k)}
        with prog.context as q:
            def look(kind, key):   # the reference's atoms (par_convert, parameters.py:249-275)
                return prog.params(key) if kind == "free" else q[key].par

            for op in self.operations:
                if op["op"] not in ops.__all__:
                    raise NameError("Quantum operation {} not defined!".format(op["op"]))
                gate = getattr(ops, op["op"])
                regs = [q[i] for i in op["modes"]]
                if op.get("args") or op.get("kwargs"):
                    gate(*[_resolve(a, look, funcs) for a in op.get("args", [])],
                         **{k: _resolve(a, look, funcs) for k, a in op.get("kwargs", {}).items()}) | regs  # noqa
                else:
                    (gate() if isinstance(gate, type) else gate) | regs  # noqa
        prog.run_options.update(self.run_options)
        prog.backend_options.update(self.backend_options)
        return prog

    # ------------------------------------------------------------------ serialisation
    def serialize(self, ir="blackbird"):
        return dumps(self, ir)


def _from_sympy(expr):
    """sympy expression over the reference's ``FreeParameter`` / ``MeasuredParameter`` atoms -> :class:`Parameter`
    (or a plain number): the direction ``to_blackbird`` needs (``blackbird_io.py:213-222``)."""
    import sympy

    def walk(e):
        if getattr(e, "is_Symbol", False):
            if hasattr(e, "regref"):
                return ("measured", int(e.regref.ind))
            return ("free", str(e.name))
        if e.is_Number or e.is_NumberSymbol or e == sympy.I:
            c = complex(e)
            if e.is_Integer:
                return int(e)
            return c.real if c.imag == 0 else c
        if e.is_Add or e.is_Mul:
            sym = "+" if e.is_Add else "*"
            kids = [walk(a) for a in e.args]
            if sym == "*" and kids[0] == -1 and len(kids) > 1:     # sympy writes -x as (-1)*x
                rest = kids[1]
                for k in kids[2:]:
                    rest = ("bin", "*", rest, k)
                return ("neg", rest) if isinstance(rest, tuple) else -rest
            out = kids[0]
            for k in kids[1:]:
                out = ("bin", sym, out, k)
            return out
        if e.is_Pow:
            base, ex = walk(e.args[0]), walk(e.args[1])
            if ex == 0.5:
                return ("call", "sqrt", (base,))
            return ("bin", "**", base, ex)
        if e.is_Function:
            name = {"Abs": "abs", "asin": "arcsin", "acos": "arccos", "atan": "arctan", "atan2": "arctan2",
                    "asinh": "arcsinh", "acosh": "arccosh", "atanh": "arctanh"}.get(type(e).__name__, type(e).__name__)
            if name not in _FUNCS:
                raise NotImplementedError("function %s has no Blackbird counterpart" % name)
            return ("call", name, tuple(walk(a) for a in e.args))
        raise NotImplementedError("cannot convert the symbolic parameter %r" % (e,))

    node = walk(sympy.sympify(expr))
    return Parameter(node) if isinstance(node, tuple) else node


def from_sf(prog, version="1.0"):
    """A ``strawberryfields.Program`` as a :class:`CircuitProgram` -- the operation list ``io.to_blackbird``
    builds (``blackbird_io.py:164-232``: measurement ``select`` / ``dark_counts`` as keyword arguments, symbolic
    parameters kept symbolic), so that ``save(f, from_sf(prog))`` writes what ``sf.save(f, prog)`` writes."""
    out = CircuitProgram(name=prog.name, version=version)
    if getattr(prog, "target", None) is not None:
        opts = dict(prog.run_options or {})
        opts.update(prog.backend_options or {})
        out.target = {"name": prog.target, "options": opts}

    def conv(a):
        if isinstance(a, np.ndarray) and a.dtype == object:
            return np.array([conv(x) for x in a.flat], dtype=object).reshape(a.shape)
        if hasattr(a, "free_symbols") or hasattr(a, "is_Symbol"):
            return _from_sympy(a)
        return a

    for cmd in prog.circuit:
        op = {"op": cmd.op.__class__.__name__, "modes": [r.ind for r in cmd.reg], "args": [], "kwargs": {}}
        if "Measure" in op["op"]:
            if cmd.op.select is not None:
                op["kwargs"]["select"] = cmd.op.select
            if cmd.op.p:
                op["args"] = [conv(a) for a in cmd.op.p]
            if op["op"] == "MeasureFock" and getattr(cmd.op, "dark_counts", None) is not None:
                op["kwargs"]["dark_counts"] = cmd.op.dark_counts
        else:
            op["args"] = [conv(a) for a in cmd.op.p]
        out.operations.append(op)
    return out


# ---------------------------------------------------------------------------------------- Blackbird
_TYPES = ("int", "float", "complex", "str", "bool")


def _loads_blackbird(text):
    prog = CircuitProgram()
    env = {}
    lines = text.splitlines()
    i = 0

    def strip(line):
        return line.split("#", 1)[0].rstrip()

    while i < len(lines):
        line = strip(lines[i])
        i += 1
        s = line.strip()
        if not s:
            continue
        head = s.split(None, 1)
        if head[0] == "name":
            prog.name = head[1].strip() if len(head) > 1 else None
            if prog.name == "None":
                prog.name = None
        elif head[0] == "version":
            prog.version = head[1].strip() if len(head) > 1 else "1.0"
        elif head[0] == "target":
            m = re.match(r"^target\s+([\w.\-]+)\s*(?:\((.*)\))?\s*$", s)
            if not m:
                raise ProgramSyntaxError("bad target line: %r" % s)
            opts = {}
            if m.group(2):
                _, opts = _parse_args(m.group(2), env, "=")
            prog.target = {"name": m.group(1), "options": opts}
        elif head[0] == "for":
            # for <type> <var> in <start>:<stop>[:<step>]   or   in [v0, v1, ..]; the indented block is unrolled
            m = re.match(r"^for\s+(int|float)\s+([A-Za-z_]\w*)\s+in\s+(.+)$", s)
            if not m:
                raise ProgramSyntaxError("bad for statement: %r" % s)
            spec = m.group(3).strip()
            if spec.startswith("["):
                values = list(_eval(spec, env))
            else:
                parts = [_eval(x, env) for x in spec.split(":")]
                if len(parts) not in (2, 3) or not all(isinstance(x, (int, np.integer)) for x in parts):
                    raise ProgramSyntaxError("bad loop range %r (start:stop[:step], integers)" % spec)
                values = list(range(int(parts[0]), int(parts[1]), int(parts[2]) if len(parts) == 3 else 1))
            block = []
            while i < len(lines) and (not strip(lines[i]).strip() or lines[i].startswith((" ", "\t"))):
                if strip(lines[i]).strip():
                    block.append(strip(lines[i]).strip())
                i += 1
            if not block:
                raise ProgramSyntaxError("empty for block")
            cast = int if m.group(1) == "int" else float
            for v in values:
                scope = dict(env)
                scope[m.group(2)] = cast(v)
                for stmt in block:
                    mm = _OP_LINE.match(stmt)
                    if not mm:
                        raise ProgramSyntaxError("only operations are allowed inside a for block: %r" % stmt)
                    args, kwargs = _parse_args(mm.group(2), scope, "=") if mm.group(2) and mm.group(2).strip() else ([], {})
                    prog.operations.append({"op": mm.group(1), "args": args, "kwargs": kwargs,
                                            "modes": _parse_modes(mm.group(3), scope)})
        elif head[0] == "type":
            # ``type tdm (temporal_modes=3)``: parsed so that the script can be inspected and converted by a front
            # end; ``loads`` refuses to hand a time-domain program to the Fock backend
            m = re.match(r"^type\s+([\w.\-]+)\s*(?:\((.*)\))?\s*$", s)
            if not m:
                raise ProgramSyntaxError("bad type line: %r" % s)
            opts = _parse_args(m.group(2), env, "=")[1] if m.group(2) else {}
            prog.programtype = {"name": m.group(1), "options": opts}
        elif head[0] == "include":
            raise NotImplementedError("Blackbird %r statements are not supported by the b200fock loader" % head[0])
        elif head[0] in _TYPES and len(head) > 1 and "=" in s and "|" not in s.split("=", 1)[0]:
            rest = head[1]
            if rest.startswith("array"):
                m = re.match(r"^array\s+([A-Za-z_]\w*)\s*(?:\[([^\]]*)\])?\s*=\s*(.*)$", rest)
                if not m:
                    raise ProgramSyntaxError("bad array declaration: %r" % s)
                rows = []
                if m.group(3).strip():
                    rows.append(m.group(3))
                while i < len(lines) and (lines[i].startswith((" ", "\t")) and strip(lines[i]).strip()):
                    rows.append(strip(lines[i]).strip())
                    i += 1
                data = [[_eval(x, env) for x in _split_top(r) if x.strip()] for r in rows]
                arr = np.array(data, dtype=object if _symbolic(data) else
                               {"int": int, "float": float, "complex": complex}.get(head[0], object))
                if m.group(2):
                    shape = [int(_eval(x, env)) for x in m.group(2).split(",") if x.strip()]
                    if list(arr.shape) != shape and arr.size == int(np.prod(shape)):
                        arr = arr.reshape(shape)
                    if list(arr.shape) != shape:
                        raise ProgramSyntaxError("array %s has shape %r, declared %r" % (m.group(1), arr.shape, shape))
                if prog.programtype["name"] == "tdm" and re.fullmatch(r"p\d+", m.group(1)):
                    # a looped-over array of a time-domain program: operations name it (blackbird keeps the name)
                    prog.variables[m.group(1)] = np.atleast_2d(arr)
                    env[m.group(1)] = Parameter.free(m.group(1))
                else:
                    env[m.group(1)] = arr
            else:
                m = re.match(r"^([A-Za-z_]\w*)\s*=\s*(.+)$", rest)
                if not m:
                    raise ProgramSyntaxError("bad variable declaration: %r" % s)
                val = _eval(m.group(2), env)
                env[m.group(1)] = {"int": int, "float": float, "complex": complex, "str": str, "bool": bool}[head[0]](val)
        else:
            m = _OP_LINE.match(s)
            if not m:
                raise ProgramSyntaxError("cannot parse line %d: %r" % (i, s))
            args, kwargs = _parse_args(m.group(2), env, "=") if m.group(2) and m.group(2).strip() else ([], {})
            if m.group(2) is None:      # written without parentheses (``Vac | 0``): an instance, not a call, in the
                prog._bare.add(len(prog.operations))   # reference front end (blackbird_io.py:63-77)
            prog.operations.append({"op": m.group(1), "args": args, "kwargs": kwargs, "modes": _parse_modes(m.group(3))})
    return prog


def _fmt(v):
    if isinstance(v, Parameter):
        return v.text
    if isinstance(v, np.ndarray):
        return "[" + ", ".join(_fmt(x) for x in v) + "]"
    if isinstance(v, (list, tuple)):
        return "[" + ", ".join(_fmt(x) for x in v) + "]"
    if isinstance(v, (complex, np.complexfloating)):
        return "%r%s%rj" % (float(v.real), "+" if v.imag >= 0 else "-", abs(float(v.imag)))
    if isinstance(v, (bool, np.bool_)) or v is None:
        return str(v)
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


def _dumps_blackbird(prog):
    out = ["name %s" % prog.name, "version %s" % (prog.version or "1.0")]
    if prog.target.get("name"):
        opts = ", ".join("%s=%s" % (k, _fmt(v)) for k, v in prog.target.get("options", {}).items())
        out.append("target %s%s" % (prog.target["name"], " (%s)" % opts if opts else ""))
    out.append("")
    arrays = 0
    body = []
    for op in prog.operations:
        parts = []
        for a in op.get("args", []):
            if isinstance(a, np.ndarray) and a.ndim == 2:
                name = "A%d" % arrays
                arrays += 1
                kind = "complex" if np.iscomplexobj(a) else "float"
                out.append("%s array %s[%d, %d] =" % (kind, name, a.shape[0], a.shape[1]))
                out.extend("    " + ", ".join(_fmt(x) for x in row) for row in a)
                out.append("")
                parts.append(name)
            else:
                parts.append(_fmt(a))
        parts += ["%s=%s" % (k, _fmt(v)) for k, v in op.get("kwargs", {}).items()]
        modes = op["modes"]
        body.append("%s(%s) | %s" % (op["op"], ", ".join(parts), modes[0] if len(modes) == 1 else "[%s]" % ", ".join(map(str, modes))))
    return "\n".join(out + body) + "\n"


# ---------------------------------------------------------------------------------------------- XIR
def _loads_xir(text):
    text = re.sub(r"//[^\n]*", "", text)
    prog = CircuitProgram(version="0.1.0")
    env = {_XIR_NAMES: True}
    for kind in ("options", "constants"):
        m = re.search(r"\b%s\s*:(.*?)\bend\s*;" % kind, text, re.S)
        if m:
            for entry in m.group(1).split(";"):
                if ":" in entry:
                    k, v = entry.split(":", 1)
                    try:
                        val = _eval(v, env)
                    except (NameError, ProgramSyntaxError):
                        if kind != "options":
                            raise
                        val = v.strip().strip('"')  # option values may be bare names (_name_: my_program)
                    if _symbolic(val):
                        if kind != "options":
                            raise NameError("constant %s refers to an undefined name: %s" % (k.strip(), v.strip()))
                        val = v.strip().strip('"')
                    if kind == "options":
                        prog.options[k.strip()] = val
                    elif str(prog.options.get("_type_", "")) == "tdm" and re.fullmatch(r"p\d+", k.strip()):
                        prog.variables[k.strip()] = val      # looped-over array: statements keep the name
                    else:
                        env[k.strip()] = val
            text = text[:m.start()] + text[m.end():]
    if "_name_" in prog.options:
        prog.name = prog.options.pop("_name_")
    if str(prog.options.get("_type_", "")) == "tdm":
        prog.programtype = {"name": "tdm", "options": {"N": prog.options.get("N")}}

    # gate definitions: ``gate Name(params)[wires]: <statements> end;`` (params and wires optional; wires default
    # to the labels of the body in sorted order).  Expanded in place at every use, as the reference does through
    # ``xir.get_expanded_statements`` (xir_io.py:89).
    defs = {}

    def take_def(m):
        body = [" ".join(x.split()) for x in m.group(4).split(";") if x.strip()]
        params = [x.strip() for x in m.group(2).split(",")] if m.group(2) and m.group(2).strip() else []
        wires = [x.strip() for x in m.group(3).split(",")] if m.group(3) and m.group(3).strip() else None
        defs[m.group(1)] = (params, wires, body)
        return ""

    text = re.sub(r"\bgate\s+([A-Za-z_]\w*)\s*(?:\(([^)]*)\))?\s*(?:\[([^\]]*)\])?\s*:(.*?)\bend\s*;", take_def, text,
                  flags=re.S)

    def wire_labels(text_modes):
        t = text_modes.strip()
        if t and t[0] in "[(":
            t = t[1:-1]
        return [x.strip() for x in _split_top(t) if x.strip()]

    def emit(stmt, scope, wire_map, depth=0):
        if depth > 32:
            raise ProgramSyntaxError("gate definitions nest deeper than 32 levels (recursive definition?)")
        m = _OP_LINE.match(stmt)
        if not m:
            raise ProgramSyntaxError("cannot parse XIR statement %r" % stmt)
        args, kwargs = _parse_args(m.group(2), scope, ":") if m.group(2) and m.group(2).strip() else ([], {})
        labels = wire_labels(m.group(3))
        if wire_map is not None:
            missing = [w for w in labels if w not in wire_map]
            if missing:
                raise ProgramSyntaxError("wire(s) %s are not wires of the gate definition" % ", ".join(missing))
            labels = [str(wire_map[w]) for w in labels]
        modes = _parse_modes("[" + ", ".join(labels) + "]")
        if m.group(1) in defs:
            params, wires, body = defs[m.group(1)]
            if kwargs:
                bound = dict(zip(params, args))
                bound.update(kwargs)
            else:
                bound = dict(zip(params, args))
            if len(args) > len(params) or sorted(bound) != sorted(params):
                raise ProgramSyntaxError("gate %s takes parameters (%s)" % (m.group(1), ", ".join(params)))
            if wires is None:
                seen = []
                for b in body:
                    mb = _OP_LINE.match(b)
                    if mb:
                        seen += [w for w in wire_labels(mb.group(3)) if w not in seen]
                wires = sorted(seen, key=lambda w: (0, int(w)) if w.isdigit() else (1, w))
            if len(wires) != len(modes):
                raise ProgramSyntaxError("gate %s acts on %d wire(s), applied to %d" % (m.group(1), len(wires), len(modes)))
            inner_scope = dict(env)
            inner_scope.update(bound)
            for b in body:
                emit(b, inner_scope, dict(zip(wires, modes)), depth + 1)
            return
        args = [np.array(a) if isinstance(a, list) else a for a in args]
        prog.operations.append({"op": m.group(1), "args": args, "kwargs": kwargs, "modes": modes})

    for stmt in text.split(";"):
        s = " ".join(stmt.split())
        if not s or s.split()[0] in ("use", "gate", "out", "func", "obs"):
            continue
        emit(s, env, None)
    return prog


def _fmt_xir(v):
    """XIR writes complex numbers the way Python prints them, in parentheses (test_io_xir.py:68-79), and free
    parameters as bare names (xir_io.py:287-303)"""
    if isinstance(v, Parameter):
        if v.measured_modes:
            raise NotImplementedError("measured parameters are written as Blackbird only")
        return re.sub(r"\{([A-Za-z_]\w*)\}", r"\1", v.text)
    if isinstance(v, (np.ndarray, list, tuple)):
        return "[" + ", ".join(_fmt_xir(x) for x in v) + "]"
    if isinstance(v, (complex, np.complexfloating)):
        return str(complex(v))
    return _fmt(v)


def _dumps_xir(prog):
    out = []
    opts = dict(prog.options)
    if prog.name:
        opts = {"_name_": prog.name, **opts}
    if opts:
        out.append("options:")
        out.extend("    %s: %s;" % (k, _fmt_xir(v)) for k, v in opts.items())
        out.append("end;")
        out.append("")
    for op in prog.operations:
        parts = [_fmt_xir(a) for a in op.get("args", [])] + ["%s: %s" % (k, _fmt_xir(v)) for k, v in op.get("kwargs", {}).items()]
        out.append("%s%s | [%s];" % (op["op"], "(%s)" % ", ".join(parts) if parts else "", ", ".join(map(str, op["modes"]))))
    return "\n".join(out)      # no trailing newline, like xir.Program.serialize()


# ------------------------------------------------------------------------------------ public API (io/__init__.py)
def _refuse_time_domain(prog):
    if prog.programtype["name"] == "tdm":
        raise NotImplementedError("time-domain (tdm) programs run on the reference's TDMProgram / hardware path, "
                                  "not on the Fock backend (DESIGN section 8)")
    if prog.programtype["name"] is not None:
        raise NotImplementedError("Blackbird program type %r is not supported" % prog.programtype["name"])


def loads(s, ir="blackbird"):
    """Load a circuit from a string (``sf.loads``, ``io/__init__.py:145-166``)."""
    if ir == "blackbird":
        prog = _loads_blackbird(s)
        _refuse_time_domain(prog)
        if not prog.operations:     # io/__init__.py:50-53: the number of modes of an empty program is unknown
            raise ValueError("Blackbird program contains no quantum operations!")
        return prog
    if ir == "xir":
        prog = _loads_xir(s)
        _refuse_time_domain(prog)
        if not prog.operations:     # xir_io.py:78-82
            raise ValueError("The XIR program is empty and cannot be transformed into a Strawberry Fields program.")
        return prog
    raise ValueError(f"'{ir}' not recognized as a valid IR option. Valid options are 'xir' and 'blackbird'.")


def load(f, ir="blackbird"):
    """Load a circuit from a ``.xbb`` / ``.xir`` file or file object (``sf.load``, ``io/__init__.py:169-237``)."""
    if hasattr(f, "read"):
        return loads(f.read(), ir)
    if not isinstance(f, (str, os.PathLike)):      # io/__init__.py:226-230
        raise ValueError("file must be a string, pathlib.Path, or file pointer")
    name = os.fspath(f)
    if ir not in ("blackbird", "xir"):
        raise ValueError(f"'{ir}' not recognized as a valid IR option. Valid options are 'xir' and 'blackbird'.")
    if ir == "blackbird" and name.endswith(".xir"):
        ir = "xir"
    with open(name) as fid:
        return loads(fid.read(), ir)


def dumps(prog, ir="blackbird"):
    if ir == "blackbird":
        return _dumps_blackbird(prog)
    if ir == "xir":
        return _dumps_xir(prog)
    raise ValueError(f"'{ir}' not recognized as a valid IR option. Valid options are 'xir' and 'blackbird'.")


def save(f, prog, ir="blackbird"):
    """Write a circuit to a ``.xbb`` / ``.xir`` file (``sf.save``, ``io/__init__.py:67-142``: the extension is
    appended when missing)."""
    text = dumps(prog, ir)
    if hasattr(f, "write"):
        f.write(text)
        return
    if not isinstance(f, (str, os.PathLike)):      # io/__init__.py:134-138
        raise ValueError("file must be a string, pathlib.Path, or file pointer")
    name = os.fspath(f)
    ext = ".xbb" if ir == "blackbird" else ".xir"
    if not name.endswith(ext):
        name += ext
    with open(name, "w") as fid:
        fid.write(text)


# ------------------------------------------------------------------------------------ lowering
def clements_rectangular(U, tol=1e-6):
    """Rectangular (Clements et al. 2016) decomposition of an N x N unitary into the gate sequence the
    reference front end emits for ``Interferometer(U)`` with its default mesh (``ops.py:2655-2718``,
    ``decompositions.rectangular``): ``Rgate(phi), BSgate(theta, 0)`` pairs, one ``Rgate`` per mode, then
    ``BSgate(-theta, 0), Rgate(-phi)`` pairs -- as backend calls on modes ``0 .. N-1``.  With
    T(theta, phi) = BS(theta, 0) R_m(phi) on neighbouring modes (m, m+1), elements of U are nulled by
    T^-1 from the right (even diagonals) and T from the left (odd diagonals) until a diagonal is left.
    ``tol``: unitarity tolerance, the reference's default (``ops.py:2640``: stored scripts print 8 digits)."""
    U = np.array(U, dtype=complex)
    N = U.shape[0]
    if U.shape != (N, N) or not np.allclose(U @ U.conj().T, np.eye(N), atol=tol):
        raise ValueError("The input matrix is not unitary")
    V = U.copy()
    right, left = [], []
    for i in range(N - 1):
        for j in range(i + 1):
            if i % 2 == 0:
                r, m = N - 1 - j, i - j                     # null V[r, m] with columns (m, m+1)
                a, b = V[r, m], V[r, m + 1]
                phi = np.angle(a) - np.angle(b) if abs(a) > 0 and abs(b) > 0 else 0.0
                theta = np.arctan2(abs(a), abs(b))
                c, s, e = np.cos(theta), np.sin(theta), np.exp(-1j * phi)
                Tinv = np.array([[e * c, e * s], [-s, c]])
                V[:, [m, m + 1]] = V[:, [m, m + 1]] @ Tinv
                right.append((m, theta, phi))
            else:
                m, col = N - 2 - i + j, j                   # null V[m+1, col] with rows (m, m+1)
                a, b = V[m, col], V[m + 1, col]
                phi = np.angle(-b) - np.angle(a) if abs(a) > 0 and abs(b) > 0 else 0.0
                theta = np.arctan2(abs(b), abs(a))
                c, s, e = np.cos(theta), np.sin(theta), np.exp(1j * phi)
                T = np.array([[e * c, -s], [e * s, c]])
                V[[m, m + 1], :] = T @ V[[m, m + 1], :]
                left.append((m, theta, phi))
    if np.abs(V - np.diag(np.diag(V))).max() > 10 * tol:
        raise ValueError("rectangular decomposition did not converge")
    calls = []
    for m, theta, phi in right:
        calls.append(("rotation", float(phi), m))
        calls.append(("beamsplitter", float(theta), 0.0, m, m + 1))
    for k in range(N):
        calls.append(("rotation", float(np.angle(V[k, k])), k))
    for m, theta, phi in reversed(left):
        calls.append(("beamsplitter", float(-theta), 0.0, m, m + 1))
        calls.append(("rotation", float(-phi), m))
    return calls


def _arg(args, kwargs, i, name, default=None):
    if i < len(args):
        return args[i]
    if name in kwargs:
        return kwargs[name]
    if default is None:
        raise TypeError("missing argument %r" % name)
    return default


def lower(op, args, kwargs, modes, cutoff=None):
    """One program operation -> ``BaseFock`` calls.  Gates whose first parameter is zero are skipped, as
    ``Gate.apply`` does (``ops.py:494-509``).  ``cutoff``: only ``Catstate`` needs it (its ket is built on the
    host, ``ops.py:880-922``)."""
    a = lambda i, name, default=None: _arg(args, kwargs, i, name, default)  # noqa: E731
    m = list(modes)
    one = {"Dgate": ("displacement", ("r", "phi")), "Sgate": ("squeeze", ("r", "phi"))}
    if op in one:
        r, phi = a(0, "r"), a(1, "phi", 0.0)
        return [] if r == 0 else [(one[op][0], float(r), float(phi), m[0])]
    if op == "Rgate":
        t = a(0, "theta")
        return [] if t == 0 else [("rotation", float(t), m[0])]
    if op == "Kgate":
        k = a(0, "kappa")
        return [] if k == 0 else [("kerr_interaction", float(k), m[0])]
    if op == "Vgate":
        g = a(0, "gamma")
        return [] if g == 0 else [("cubic_phase", float(g), m[0])]
    if op == "Fouriergate":
        return [("rotation", math.pi / 2, m[0])]
    if op == "Xgate":
        x = a(0, "x")
        return [] if x == 0 else [("displacement", abs(float(x)) / math.sqrt(2 * HBAR), 0.0 if x > 0 else math.pi, m[0])]
    if op == "Zgate":
        p = a(0, "p")
        return [] if p == 0 else [("displacement", abs(float(p)) / math.sqrt(2 * HBAR),
                                   math.pi / 2 if p > 0 else -math.pi / 2, m[0])]
    if op == "Pgate":  # ops.py:1726-1732
        s = float(a(0, "s"))
        if s == 0:
            return []
        temp = s / 2
        r = math.acosh(math.sqrt(1 + temp ** 2))
        theta = math.atan(temp)
        phi = -math.pi / 2 * np.sign(temp) - theta
        return [("squeeze", r, float(phi), m[0]), ("rotation", theta, m[0])]
    if op == "BSgate":
        return [("beamsplitter", float(a(0, "theta", math.pi / 4)), float(a(1, "phi", 0.0)), m[0], m[1])] \
            if a(0, "theta", math.pi / 4) != 0 else []
    if op == "MZgate":
        return [("mzgate", float(a(0, "phi_in")), float(a(1, "phi_ex")), m[0], m[1])]
    if op == "sMZgate":  # ops.py:2023-2030
        bs = ("beamsplitter", math.pi / 4, math.pi / 2, m[0], m[1])
        rots = [("rotation", float(a(1, "phi_ex")) - math.pi / 2, m[1]), ("rotation", float(a(0, "phi_in")) - math.pi / 2, m[0])]
        return [bs] + [c for c in rots if c[1] != 0] + [bs]
    if op == "S2gate":
        r = a(0, "r")
        return [] if r == 0 else [("two_mode_squeeze", float(r), float(a(1, "phi", 0.0)), m[0], m[1])]
    if op == "CKgate":
        k = a(0, "kappa")
        return [] if k == 0 else [("cross_kerr_interaction", float(k), m[0], m[1])]
    if op == "CXgate":  # ops.py:2149-2158
        s = float(a(0, "s", 1.0))
        if s == 0:
            return []
        r = math.asinh(-s / 2)
        theta = 0.5 * math.atan2(-1.0 / math.cosh(r), -math.tanh(r))
        return [("beamsplitter", theta, 0.0, m[0], m[1]), ("squeeze", r, 0.0, m[0]), ("squeeze", -r, 0.0, m[1]),
                ("beamsplitter", theta + math.pi / 2, 0.0, m[0], m[1])]
    if op == "CZgate":  # ops.py:2209-2216
        s = float(a(0, "s", 1.0))
        if s == 0:
            return []
        return [("rotation", -math.pi / 2, m[1])] + lower("CXgate", [s], {}, m) + [("rotation", math.pi / 2, m[1])]
    if op == "Interferometer":
        U = np.asarray(a(0, "U"), dtype=complex)
        if U.shape != (len(m), len(m)):
            raise ValueError("Interferometer matrix is %r for %d modes" % (U.shape, len(m)))
        if kwargs.get("mesh", "rectangular") != "rectangular":
            raise NotImplementedError("only the rectangular mesh is decomposed by the b200fock loader")
        out = []
        for c in clements_rectangular(U):
            if c[1] == 0:
                continue
            out.append(c[:-1] + (m[c[-1]],) if c[0] == "rotation" else c[:-2] + (m[c[-2]], m[c[-1]]))
        return out
    if op == "LossChannel":
        return [("loss", float(a(0, "T")), m[0])]
    # ---- preparations
    if op in ("Vacuum", "Vac"):
        return [("prepare_vacuum_state", x) for x in m]
    if op == "Fock":
        return [("prepare_fock_state", int(a(0, "n", 0) if not (args or kwargs) else a(0, "n")), m[0])]
    if op == "Coherent":
        return [("prepare_coherent_state", float(a(0, "r", 0.0)), float(a(1, "phi", 0.0)), m[0])]
    if op == "Squeezed":
        return [("prepare_squeezed_state", float(a(0, "r", 0.0)), float(a(1, "p", 0.0)), m[0])]
    if op == "DisplacedSqueezed":
        return [("prepare_displaced_squeezed_state", float(a(0, "r_d", 0.0)), float(a(1, "phi_d", 0.0)),
                 float(a(2, "r_s", 0.0)), float(a(3, "phi_s", 0.0)), m[0])]
    if op == "Thermal":
        return [("prepare_thermal_state", float(a(0, "n", 0.0)), m[0])]
    if op == "Catstate":  # ops.py:880-922: (|alpha> + e^{i pi p} |-alpha>) / N as a host-built ket
        if cutoff is None:
            raise ValueError("Catstate needs the cutoff dimension: calls(cutoff_dim=..) or run()")
        alpha = float(a(0, "a", 0.0)) * np.exp(1j * float(a(1, "phi", 0.0)))
        theta = math.pi * float(a(2, "p", 0))
        l = np.arange(int(cutoff))
        fact = np.sqrt(np.array([math.factorial(int(x)) for x in l], dtype=float))
        temp = math.exp(-0.5 * abs(alpha) ** 2)
        N = temp / math.sqrt(2 * (1 + math.cos(theta) * temp ** 4))
        ket = (alpha ** l / fact + np.exp(1j * theta) * (-alpha) ** l / fact) * N
        return [("prepare_ket_state", ket, m)]
    if op == "GKP":  # ops.py:959-967 -> backend.prepare_gkp(state, epsilon, ampl_cutoff, representation, shape, mode)
        state = a(0, "state", [0, 0])
        return [("prepare_gkp", [float(x) for x in state], float(a(1, "epsilon", 0.2)), float(a(2, "ampl_cutoff", 1e-12)),
                 a(3, "representation", "real"), a(4, "shape", "square"), {"mode": m[0]})]
    if op == "Ket":
        return [("prepare_ket_state", np.asarray(a(0, "state")), m)]
    if op == "DensityMatrix":
        return [("prepare_dm_state", np.asarray(a(0, "state")), m)]
    # ---- measurements
    if op == "MeasureFock":
        sel = kwargs.get("select")
        if sel is not None and not isinstance(sel, (list, tuple)):
            sel = [sel]
        return [("measure_fock", m, {"select": sel})] if sel is not None else [("measure_fock", m)]
    if op in ("MeasureHomodyne", "MeasureX", "MeasureP", "MeasureHD"):
        phi = {"MeasureX": 0.0, "MeasureP": math.pi / 2}.get(op)
        if phi is None:
            phi = float(a(0, "phi", 0.0))
        sel = kwargs.get("select")
        return [("measure_homodyne", phi, m[0], {"select": sel})]
    raise NotImplementedError("operation %s is not supported by the Fock backend loader" % op)


# ------------------------------------------------------------------------------------ state checkpoints
CHECKPOINT_VERSION = 1


def save_state(f, state):
    """Write the ket / density matrix of a state object (``B200FockState``, or anything with ``data``,
    ``is_pure``, ``num_modes``, ``cutoff_dim``) to an ``.npz`` checkpoint."""
    data = np.asarray(state.data)
    np.savez_compressed(f, data=data, pure=np.array(bool(state.is_pure)), num_modes=np.array(int(state.num_modes)),
                        cutoff_dim=np.array(int(state.cutoff_dim)), version=np.array(CHECKPOINT_VERSION),
                        layout=np.array("ket [D]*n | dm (ket_0, bra_0, ket_1, ...) [D]*2n, C order, complex128"))


def load_state(f, backend=None):
    """Read a checkpoint; with ``backend`` (already in a circuit of the same size) the state is prepared on
    it (``prepare_ket_state`` / ``prepare_dm_state`` on all modes).  Returns ``(data, pure)``."""
    with np.load(f, allow_pickle=False) as z:
        if int(z["version"]) != CHECKPOINT_VERSION:
            raise ValueError("unsupported checkpoint version %d" % int(z["version"]))
        data, pure, n, D = z["data"], bool(z["pure"]), int(z["num_modes"]), int(z["cutoff_dim"])
    if backend is not None:
        if backend.get_cutoff_dim() != D or len(backend.get_modes()) != n:
            raise ValueError("checkpoint holds %d modes at cutoff %d; the backend's circuit differs" % (n, D))
        # a batched state ([B, ..]) goes back in one call: the leading axis is the batch (tfbackend/circuit.py:343-396)
        (backend.prepare_ket_state if pure else backend.prepare_dm_state)(data, list(range(n)))
    return data, pure
