"""The C-ABI library loads and exports every symbol include/*.h declares; struct layouts match the ctypes mirror."""
import ctypes as C
import os
import re

from wave_tracer_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(wt(?:gpu|host)_[a-z_0-9]+)\s*\(", src))


def test_every_declared_symbol_is_exported():
    """wtgpu.h -> libwt_b200.so (the CUDA library), wthost.h -> libwt_host.so (host-only scene preparation)."""
    L, H = A.lib(), A.host_lib()
    gpu, host = _declared("wtgpu.h"), _declared("wthost.h")
    assert len(gpu) >= 14 and len(host) >= 6
    for n in sorted(gpu):
        assert hasattr(L, n), f"{n} declared in include/wtgpu.h but not exported by libwt_b200.so"
    for n in sorted(host):
        assert hasattr(H, n), f"{n} declared in include/wthost.h but not exported by libwt_host.so"
    assert set(A.EXPORTED_SYMBOLS) <= gpu and set(A.HOST_EXPORTED_SYMBOLS) <= host


def test_host_library_is_host_only():
    """libwt_host.so (what the CPU legs of bench.py build their scene tables with) links neither the CUDA runtime nor the CUDA library, and
    building a scene does not load libwt_b200.so."""
    import subprocess, sys
    out = subprocess.run(["ldd", A.HOST_LIB_PATH], capture_output=True, text=True).stdout
    assert "cudart" not in out and "libcuda" not in out and "libwt_b200" not in out, out
    code = ("import sys; sys.path.insert(0, %r)\nfrom wave_tracer_b200 import scenes, _abi\nb = scenes.double_slits(res=32, spp=1).build()\n"
            "assert b.desc.n_tris == 10 and _abi._lib is None\nassert 'libwt_b200' not in open('/proc/self/maps').read()\nprint('ok')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1500:]


def test_struct_layouts_match():
    L = A.lib()
    for i, s in enumerate(A.ABI_STRUCTS):
        assert C.sizeof(s) == L.wtgpu_debug_sizeof(i), (s.__name__, C.sizeof(s), L.wtgpu_debug_sizeof(i))
    assert C.sizeof(A.Node) == 256 and C.sizeof(A.Tri) == 48 and C.sizeof(A.TriMeta) == 32 and C.sizeof(A.Edge) == 96


def test_no_gpu_fails_loudly():
    """Without a device the product path reports an error -- it never falls back to a CPU implementation."""
    L = A.lib()
    if L.wtgpu_device_count() > 0:
        return
    from wave_tracer_b200 import scenes, GpuScene
    import pytest
    built = scenes.double_slits(res=32, spp=1, with_directional=False).build()
    with pytest.raises(RuntimeError):
        GpuScene(built, 0)


def test_product_never_imports_oracle():
    """Static guard: nothing under wave_tracer_b200/ may reference oracle/."""
    pkg = os.path.join(ROOT, "wave_tracer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("oracle/.", "") or f == "__init__.py" and False, f"{f} references oracle/"
                assert "liboracle" not in txt and "_oracle" not in txt, f"{f} references the oracle"
