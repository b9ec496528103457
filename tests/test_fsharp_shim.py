"""Static consistency of the F# P/Invoke twin (integration/Backend.Cuda.fs) with the C ABI (include/diffsharp_cuda.h).

The shim cannot be compiled here (no .NET in this image), so what can be checked without a compiler is checked: every
`extern` it declares names an entry point the header declares, with the same number of parameters and compatible
parameter kinds (handle / pointer-out / scalar), and its explicit MapOp -> op-code table equals the ordinals of the
reference's MapOp union (Backend.fs:42-70) that the Python twin and the header use."""
import os
import re

from sigma.diffsharp_b200.backend import MapOp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FS = os.path.join(ROOT, "integration", "Backend.Cuda.fs")
HDR = os.path.join(ROOT, "include", "diffsharp_cuda.h")


def _split_params(text):
    text = text.strip()
    if not text or text == "void":
        return []
    return [p.strip() for p in text.split(",") if p.strip()]


def header_prototypes():
    """name -> list of C parameter declarations (typed macro expanded for s / d)."""
    src = re.sub(r"/\*.*?\*/", " ", open(HDR).read(), flags=re.S)
    src = src.replace("\\\n", " ")
    protos = {}
    m = re.search(r"#define DSC_DECLARE_TYPED\(P, T\)(.*?)\n\s*\n", src, flags=re.S)
    body = m.group(1)
    for name, params in re.findall(r"dsc_##P##(\w+)\s*\(([^)]*)\)", body):
        for p, t in (("s", "float"), ("d", "double")):
            protos[f"dsc_{p}{name}"] = [re.sub(r"\bT\b", t, q) for q in _split_params(params)]
    rest = src.replace(body, " ")
    for name, params in re.findall(r"DSC_API\s+[\w\s\*]+?\b(dsc_\w+)\s*\(([^)]*)\)\s*;", rest):
        protos[name] = _split_params(params)
    return protos


def fsharp_externs():
    src = open(FS).read()
    out = {}
    for name, params in re.findall(r"extern\s+\w+\s+(dsc_\w+)\s*\(([^)]*)\)", src):
        out[name] = _split_params(params)
    return out


def _kind_c(decl):
    d = decl.strip()
    if "*" in d or "[" in d:
        return "pointer"
    if re.match(r"(const\s+)?dsc_buf_t\b", d) or re.match(r"(const\s+)?(dsc_ctx_t|dsc_graph_t|dsc_event_t)\b", d):
        return "handle"
    return "scalar"


def _kind_fs(decl):
    d = decl.strip()
    ty = d.split()[0]
    if "&" in ty or ty.endswith("[]") or ty in ("nativeint&",):
        return "pointer"
    if ty in ("uint64", "nativeint"):
        return "handle_or_pointer"   # handles are uint64 / nativeint; raw host pointers are nativeint too
    return "scalar"


def test_every_fsharp_extern_is_declared_with_the_same_arity():
    protos, ext = header_prototypes(), fsharp_externs()
    assert len(ext) >= 60, "the shim binds the whole typed surface"
    for name, fparams in ext.items():
        assert name in protos, f"{name}: bound by the F# shim but not declared in include/diffsharp_cuda.h"
        cparams = protos[name]
        assert len(fparams) == len(cparams), f"{name}: F# declares {len(fparams)} parameters, the header {len(cparams)}: {fparams} vs {cparams}"
        for fp, cp in zip(fparams, cparams):
            kf, kc = _kind_fs(fp), _kind_c(cp)
            if kf == "handle_or_pointer":
                assert kc in ("handle", "pointer"), f"{name}: `{fp}` against `{cp}`"
            else:
                assert kf == kc, f"{name}: `{fp}` ({kf}) against `{cp}` ({kc})"


def test_both_element_types_are_bound_symmetrically():
    ext = fsharp_externs()
    s = {n[5:] for n in ext if n.startswith("dsc_s") and "dsc_d" + n[5:] in header_prototypes()}
    d = {n[5:] for n in ext if n.startswith("dsc_d") and "dsc_s" + n[5:] in header_prototypes()}
    typed_s = {n for n in s if f"dsc_d{n}" in ext}
    assert typed_s == {n for n in d if f"dsc_s{n}" in ext} and len(typed_s) >= 25


def test_mapop_codes_match_the_reference_ordinals():
    src = open(FS).read()
    table = {m.group(1): int(m.group(2)) for m in re.finditer(r"MapOp\.(\w+)\s*->\s*(\d+)", src)}
    assert len(table) == 28
    for name, code in table.items():
        assert int(getattr(MapOp, name)) == code, name
    assert sorted(table.values()) == list(range(28))
