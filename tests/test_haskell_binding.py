"""The Haskell side under haskell/ cannot be compiled here (no GHC), so its contact surface with the C-ABI is checked
statically: every `foreign import ccall "name"` must name a symbol that the built libraries export and that
include/glome_cuda.h declares, with as many arguments as the Haskell type has; the byte offsets the binding pokes / peeks
must be the ones GLOME_LAYOUT_ASSERT pins in the header; and the patch must still apply to the reference tree when that
tree is present (it is not on the GPU box)."""
import os
import re
import shutil
import subprocess
import tempfile

import pytest

from glome_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HS = [os.path.join(ROOT, "haskell", "Data", "Glome", "CUDA.hs"), os.path.join(ROOT, "haskell", "Data", "Glome", "CUDA", "Flat.hs")]
HEADER = open(os.path.join(ROOT, "include", "glome_cuda.h")).read()


def foreign_imports():
    out = []
    for path in HS:
        src = open(path).read()
        for m in re.finditer(r'foreign import ccall (?:safe|unsafe)\s+"(&?)(\w+)"\s+\w+\s*::\s*(.*)', src):
            addr, name, ty = m.group(1) == "&", m.group(2), m.group(3).strip()
            out.append((os.path.basename(path), name, addr, ty))
    return out


def c_prototypes():
    flat = re.sub(r"/\*.*?\*/", " ", HEADER, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|void|const char\*)\s+(glome_\w+)\s*\(([^;{]*?)\)\s*;", flat, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        protos[m.group(1)] = n
    return protos


def hs_arity(ty):
    """number of arguments of a Haskell function type `a -> b -> IO r` (no higher-order arguments in this binding)"""
    depth, parts, cur = 0, [], ""
    i = 0
    while i < len(ty):
        c = ty[i]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
        if depth == 0 and ty.startswith("->", i):
            parts.append(cur.strip())
            cur = ""
            i += 2
            continue
        cur += c
        i += 1
    parts.append(cur.strip())
    return len(parts) - 1


def test_every_foreign_import_names_a_declared_and_exported_symbol_with_the_same_arity():
    imports, protos = foreign_imports(), c_prototypes()
    assert len(imports) >= 50
    exported = set()
    for lib in (L.LIB_PATH, L.HOST_LIB_PATH):
        out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
        exported |= {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    for fname, name, addr, ty in imports:
        assert name in protos, "%s: %s is not declared in include/glome_cuda.h" % (fname, name)
        assert name in exported, "%s: %s is not exported by the libraries" % (fname, name)
        if addr:
            continue  # a FunPtr import (the finalizer)
        assert hs_arity(ty) == protos[name], "%s: %s has %d arguments in Haskell, %d in C" % (fname, name, hs_arity(ty), protos[name])


def test_struct_offsets_used_by_the_binding_are_the_asserted_ones():
    cuda_hs = open(HS[0]).read()
    # pokeOpts: mode 0, blocksize 4, recurs 8, thresholds 16..40
    for off in (0, 4, 8, 16, 24, 32, 40):
        assert re.search(r"pokeByteOff p %d\b" % off, cuda_hs), off
    for field, off in (("mode", 0), ("blocksize", 4), ("recurs", 8), ("thresholds", 16)):
        assert re.search(r"offsetof\(GlomeRenderOpts, %s\) == %d\b" % (field, off), HEADER), field
    # CudaHit: hit 56, prim 60, sub 64, ntex 68, ntag 72, tex 80, tag 112, size 144
    for field, off in (("hit", 56), ("prim", 60), ("sub", 64), ("ntex", 68), ("ntag", 72), ("tex", 80), ("tag", 112)):
        assert re.search(r"offsetof\(GlomeHit, %s\) == %d\b" % (field, off), HEADER), field
        assert re.search(r"(peekByteOff p %d\b|peekByteOff p \(%d\+)" % (off, off), cuda_hs), (field, off)
    assert "sizeOf _ = 144" in cuda_hs and "sizeof(GlomeHit) == 144" in HEADER
    assert "allocaBytes 256" in cuda_hs and "sizeof(GlomeFlatScene) <= 256" in HEADER
    assert "allocaBytes 64" in cuda_hs and "sizeof(GlomeRenderOpts) == 64" in HEADER


def test_every_solid_of_the_reference_gets_a_flatten_instance():
    patch = open(os.path.join(ROOT, "haskell", "glometrace-cuda.patch")).read()
    added = "\n".join(ln[1:] for ln in patch.splitlines() if ln.startswith("+") and not ln.startswith("+++"))
    for solid in ("SolidItem s", "Void", "Instance s xfm", "Sphere c r _", "Triangle p1 p2 p3", "TriangleNorm p1 p2 p3 n1 n2 n3",
                  "Box bb", "Plane n off", "Disc pos n rsqr", "Cylinder r h1 h2", "Cone r c1 c2 h", "Difference a b useatex",
                  "Intersection ss", "Tag s tag", "Tex s tex", "TexD s _ desc", "NoShadow s", "OnlyShadow s", "Bound sa sb",
                  "InnerBound sa sb"):
        assert re.search(r"flatten fb \(?%s\)?" % re.escape(solid), added), solid
    assert "flatten fb xs" in added and "flatten = flatten_bih" in added and "flatten = flatten_mesh" in added
    assert "emitBihPrebuilt" in added and "emitMeshPrebuilt" in added  # the trees the constructors built, not rebuilt


@pytest.mark.skipif(not os.path.isdir("/root/reference/GlomeTrace"), reason="the reference tree is not on this machine")
def test_patch_applies_to_the_reference_tree():
    tmp = tempfile.mkdtemp()
    try:
        shutil.copytree("/root/reference/GlomeTrace", os.path.join(tmp, "GlomeTrace"))
        r = subprocess.run(["patch", "-p1", "--dry-run", "-i", os.path.join(ROOT, "haskell", "glometrace-cuda.patch")], cwd=tmp,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
