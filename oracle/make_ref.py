"""Build ``oracle/_ref/libspconv_ref.so``: the REFERENCE's own CPU rulebook + gather/scatter code,
compiled from the sources where they lie under /root/reference.  TEST INFRASTRUCTURE ONLY.

The reference's CPU implementation of this path is plain C++ held in Python f-strings that the
``pccm`` code generator assembles into classes:

* ``spconv/csrc/sparse/indices.py:77-269``    ``ConvOutLocIter``   (coordinate iterator)
* ``spconv/csrc/sparse/indices.py:1621-1778`` ``SparseConvIndicesCPU`` (``generate_subm_conv_inds``,
  ``generate_conv_inds``: the hash-map rulebook loops)
* ``spconv/csrc/sparse/gather.py:30-86``      ``GatherCPU`` (``gather`` / ``scatter_add``)
* ``spconv/csrc/sparse/maxpool.py:590-700``   ``IndiceMaxPoolCPU`` (``forward`` / ``backward`` /
  ``global_pool_rearrange``)
* ``spconv/csrc/sparse/pointops.py:42-88,589-695`` ``Point2VoxelCommon::calc_meta_data`` and
  ``Point2VoxelCPU::point_to_voxel_static`` (+ the ``empty_mean`` variant)

``pccm`` / ``cumm`` / ``ccimport`` are not installable here (no network), so the reference's own
build cannot run.  This script instead *executes the reference's generator methods* against a
minimal stand-in for the ``pccm`` API (a ``FunctionCode`` that records ``arg`` / ``raw`` / ``ret``
calls; ``codeops.unpack`` / ``dispatch_ints``), collects the emitted C++ text verbatim, wraps every
class in a namespace per ``ndim`` and compiles it with g++ against ``oracle/ref_shim.h`` -- a
~150-line header restating the few ``tv::`` / cumm types the text uses (``tv::Tensor``,
``tv::array``, ``TensorGeneric`` row-major layout, ``ConvProblem``).  cumm itself is an
un-vendored dependency (``cumm>=0.7.11,<0.8.0``, pyproject.toml); the shim follows its published
semantics, everything else is the reference's text.

Outputs go to ``oracle/_ref/`` only (git-ignored; it travels to the GPU box like any built .so).
No reference source is copied into the repository: the generated header lives next to the .so.

    python oracle/make_ref.py            # build (no-op when up to date)
    python oracle/make_ref.py --force
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import subprocess
import sys
import types
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SPCONV_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
GEN_HDR = os.path.join(OUT_DIR, "spconv_ref_gen.h")
LIB = os.path.join(OUT_DIR, "libspconv_ref.so")
SHIM = os.path.join(HERE, "ref_shim.h")
CAPI = os.path.join(HERE, "ref_capi.cpp")
REF_FILES = ["spconv/csrc/sparse/indices.py", "spconv/csrc/sparse/gather.py", "spconv/csrc/sparse/maxpool.py",
             "spconv/csrc/sparse/pointops.py"]


def reference_available() -> bool:
    return all(os.path.exists(os.path.join(REF_ROOT, f)) for f in REF_FILES)


# ------------------------------------------------------------------------------ pccm stand-in
class FunctionCode:
    """Records what a generator method emits (the subset of pccm.FunctionCode these files use)."""

    def __init__(self):
        self.args: List[tuple] = []
        self.targs: List[str] = []
        self.body: List[str] = []
        self.ret_type: Optional[str] = None
        self.inits: List[tuple] = []

    def arg(self, names, ctype, default=None, **_):
        for n in names.split(","):
            self.args.append((n.strip(), ctype, default))
        return self

    def targ(self, name):
        self.targs.append(f"typename {name}")
        return self

    def nontype_targ(self, name, ctype):
        self.targs.append(f"{ctype} {name}")
        return self

    def raw(self, text):
        self.body.append(text)
        return self

    def ctor_init(self, name, value):
        self.inits.append((name, value))
        return self

    def ret(self, ctype, **_):
        self.ret_type = ctype
        return self


class _Any:
    """Absorbs every attribute access / call the class bodies make on pccm / cumm objects."""

    def __init__(self, name="any"):
        self._name = name

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Any(f"{self._name}.{item}")

    def __call__(self, *a, **k):
        # decorator use: @pccm.static_function  /  @pccm.member_function(header_only=True, ...)
        if len(a) == 1 and callable(a[0]) and not isinstance(a[0], _Any) and not k:
            return a[0]
        return _Any(self._name + "()")

    def __str__(self):
        return self._name


class _PccmClass:
    def __init__(self, *a, **k):
        self._members = []

    def add_dependency(self, *a, **k): pass
    def add_param_class(self, *a, **k): pass
    def add_include(self, *a, **k): pass
    def add_member(self, name, ctype, *a, **k): self._members.append((name, ctype))
    def add_static_const(self, *a, **k): pass
    def add_enum_class(self, *a, **k): pass
    def add_pybind_member(self, *a, **k): pass

    @property
    def class_name(self):
        return type(self).__name__


class _DType:
    def __init__(self, c): self.c = c
    def __str__(self): return self.c
    def __format__(self, spec): return self.c


def _install_stubs():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__getattr__ = lambda item, _n=name: _Any(f"{_n}.{item}")   # PEP 562
        sys.modules[name] = m
        return m

    def unpack(name, rng, left="[", right="]"):   # cumm.gemm.codeops.unpack: "x[0], x[1], ..."
        return ", ".join(f"{name}{left}{i}{right}" for i in rng)

    def unpack_str(name, rng):      # "h_0, h_1, ..."
        return ", ".join(f"{name}_{i}" for i in rng)

    def dispatch_ints(code, ints, expr):   # if / else-if chain over the listed values
        for n, x in enumerate(ints):
            code.raw(("if" if n == 0 else "else if") + f" ({expr} == {x}) {{")
            yield x
            code.raw("}")

    pccm = mod("pccm", FunctionCode=FunctionCode, code=FunctionCode, ParameterizedClass=_PccmClass,
               Class=_PccmClass, literal=lambda v: repr(v),
               boolean=lambda v: "true" if v else "false")
    pccm.cuda = _Any("pccm.cuda")
    pccm.pybind = _Any("pccm.pybind")
    pccm.pybind.PybindClassMixin = type("PybindClassMixin", (), {})          # used as a base class
    dtypes = mod("cumm.dtypes", int32=_DType("int32_t"), int64=_DType("int64_t"), float32=_DType("float"))
    mod("cumm", dtypes=dtypes)
    mod("cumm.gemm")
    mod("cumm.gemm.core")
    mod("cumm.gemm.core.metaarray")
    mod("cumm.gemm.layout", TensorGeneric=lambda *a, **k: _Any("TensorGeneric"))
    mod("cumm.common")
    mod("cumm.constants", CUMM_CPU_ONLY_BUILD=True)
    codeops = mod("cumm.gemm.codeops", unpack=unpack, unpack_str=unpack_str, dispatch_ints=dispatch_ints)
    sys.modules["cumm.gemm"].codeops = codeops
    mod("cumm.conv")
    mod("cumm.conv.params")
    mod("spconv")
    mod("spconv.csrc")
    mod("spconv.csrc.sparse")
    mod("spconv.csrc.sparse.cpu_core")
    mod("spconv.csrc.utils")
    mod("spconv.csrc.utils.launch")
    mod("cumm.gemm.mask_iters")
    return dtypes


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# ------------------------------------------------------------------------------ emit C++
def _emit_fn(name: str, code: FunctionCode, static: bool, const: bool = False, ctor_of: str = "") -> str:
    tpl = f"template <{', '.join(code.targs)}>\n" if code.targs else ""
    args = ", ".join(f"{t} {n}" + (f" = {d}" if d is not None else "") for n, t, d in code.args)
    body = "\n".join(code.body)
    if ctor_of:
        inits = ", ".join(f"{n}({v})" for n, v in code.inits)
        return f"{ctor_of}({args}) : {inits} {{\n{body}\n}}\n"
    ret = code.ret_type or "void"
    return (f"{tpl}{'static ' if static else ''}{ret} {name}({args}){' const' if const else ''} "
            f"{{\n{body}\n}}\n")


def generate() -> str:
    dtypes = _install_stubs()
    ind = _load(os.path.join(REF_ROOT, REF_FILES[0]), "_ref_indices")
    gat = _load(os.path.join(REF_ROOT, REF_FILES[1]), "_ref_gather")
    mpl = _load(os.path.join(REF_ROOT, REF_FILES[2]), "spconv.csrc.sparse.maxpool")   # has a relative import
    out = ["// GENERATED by oracle/make_ref.py from the reference's own C++ text -- do not commit.\n",
           '#pragma once\n#include "../ref_shim.h"\n']
    for ndim in (1, 2, 3, 4):
        problem = types.SimpleNamespace(ndim=ndim)
        out.append(f"namespace ref_nd{ndim} {{\n")
        out.append(f"using ConvProblem = refshim::ConvProblem<{ndim}>;\n")
        for cls_name, use_i64 in (("ConvLocIter", False), ("ConvLocIter64", True)):
            it = ind.ConvOutLocIter(problem, use_i64)
            idx_t = "int64_t" if use_i64 else "int32_t"
            out.append(f"struct {cls_name} {{\n")
            out.append(f"using LayoutNPQ = refshim::TensorGeneric<{ndim + 1}, {idx_t}>;\n")
            out.append(f"using LayoutRS = refshim::TensorGeneric<{ndim}, int32_t>;\n")
            for n, t in it._members:
                out.append(f"{t} {n};\n")
            out.append(_emit_fn("", it.ctor(), False, ctor_of=cls_name))
            inc = it.increment()
            inc.ret_type = f"{cls_name}&"
            out.append(_emit_fn("operator++", inc, False))
            out.append(_emit_fn("set_filter_offset", it.set_filter_offset(), False))
            for meth in ("nhw_to_npq", "npq_to_nhw", "query_npq", "query_npq_no_stride", "query_nhw",
                         "query_nhw_out"):
                out.append(_emit_fn(meth, getattr(it, meth)(), False, const=True))
            out.append("};\n")
        cpu = ind.SparseConvIndicesCPU(problem, dtypes.int32)
        out.append("struct SparseConvIndicesCPU {\n")
        out.append(_emit_fn("generate_subm_conv_inds", cpu.generate_subm_conv_inds(), True))
        out.append(_emit_fn("generate_conv_inds", cpu.generate_conv_inds(), True))
        out.append("};\n")
        out.append(f"}}  // namespace ref_nd{ndim}\n")
    g = gat.GatherCPU()
    out.append("struct GatherCPU {\n")
    out.append(_emit_fn("gather", g.gather(), True))
    out.append(_emit_fn("scatter_add", g.scatter_add(), True))
    out.append("};\n")
    pts = _load(os.path.join(REF_ROOT, REF_FILES[3]), "_ref_pointops")
    for ndim in (2, 3):
        out.append(f"namespace ref_p2v{ndim} {{\n")
        common = pts.Point2VoxelCommon(dtypes.float32, ndim, True)
        out.append("struct Point2VoxelCommon {\n")
        out.append(_emit_fn("calc_meta_data", common.calc_meta_data(), True))
        out.append("};\n")
        cpu = pts.Point2VoxelCPU(dtypes.float32, ndim, True)
        out.append("struct Point2VoxelCPU {\n")
        out.append(_emit_fn("point_to_voxel_static", cpu.point_to_voxel_static_template(False), True))
        out.append(_emit_fn("point_to_voxel_empty_mean_static", cpu.point_to_voxel_static_template(True), True))
        out.append("};\n")
        out.append(f"}}  // namespace ref_p2v{ndim}\n")
    mp = mpl.IndiceMaxPoolCPU()
    out.append("struct IndiceMaxPoolCPU {\n")
    out.append(_emit_fn("global_pool_rearrange", mp.global_pool_rearrange(), True))
    out.append(_emit_fn("forward", mp.forward(), True))
    out.append(_emit_fn("backward", mp.backward(), True))
    out.append("};\n")
    return "".join(out)


def build(force: bool = False, verbose: bool = False) -> Optional[str]:
    """Returns the library path, or None when the reference tree is not present (GPU box: the
    prebuilt file that travelled with the snapshot is used as is)."""
    if not reference_available():
        return LIB if os.path.exists(LIB) else None
    deps = [os.path.join(REF_ROOT, f) for f in REF_FILES] + [SHIM, CAPI, os.path.abspath(__file__)]
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    saved = dict(sys.modules)
    try:
        text = generate()
    finally:                         # the stand-in modules must not leak into the caller's process
        for k in list(sys.modules):
            if k not in saved:
                del sys.modules[k]
    with open(GEN_HDR, "w") as f:
        f.write(text)
    # No -fopenmp: the reference's default (CUDA) build compiles GatherCPU without OMPLib
    # (gather.py:25-26: OpenMP only when CUMM_CPU_ONLY_BUILD), and a second OpenMP runtime next to
    # torch's own thread pool makes every parallel region ~100x slower (measured: 0.17 -> 47 ms).
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unused-variable",
           "-I", HERE, "-o", LIB, CAPI]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"g++ failed on the extracted reference code:\n{res.stderr[:6000]}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print(p if p else "reference tree not present and no prebuilt oracle/_ref library")
