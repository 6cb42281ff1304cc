"""The C-ABI library loads and exports every symbol include/orcdchomp_b200.h declares;
without a GPU every entry point refuses loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from or_cdchomp_b200 import capi


def declared_functions():
    src = open(os.path.join(ROOT, "include", "orcdchomp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ocb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_bindings_agree():
    names = declared_functions()
    assert len(names) >= 30
    assert sorted(capi.EXPORTS) == names


def test_library_exports_every_symbol():
    lib = capi.load_library()
    for name in declared_functions():
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.ocb_version()


def test_struct_sizes_match_header():
    """ctypes mirrors vs the C compiler's layout (checked by compiling a probe)."""
    import subprocess
    import tempfile
    probe = r'''
    #include <stdio.h>
    #include "orcdchomp_b200.h"
    int main(void){ printf("%zu %zu %zu %zu\n", sizeof(ocb_robot), sizeof(ocb_sdf), sizeof(ocb_params), sizeof(ocb_prim)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "p.c")
        open(c, "w").write(probe)
        exe = os.path.join(d, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(capi.OcbRobot), C.sizeof(capi.OcbSdf), C.sizeof(capi.OcbParams), C.sizeof(capi.OcbPrim)]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = capi.load_library()
    h = C.c_void_p()
    rc = lib.ocb_engine_create(0, C.byref(h))
    assert rc == capi.OCB_ERR_NODEVICE and not h.value
    assert b"no CPU path" in lib.ocb_last_error()
    from or_cdchomp_b200.engine import Engine
    with pytest.raises(capi.OcbError):
        Engine(0)


def test_product_does_not_import_oracle():
    """nothing under or_cdchomp_b200/ may reference the oracle."""
    pkg = os.path.join(ROOT, "or_cdchomp_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "import oracle" not in text, f


def test_libcd_compat_library_exports_and_refuses_without_gpu():
    """include/libcd_b200.h: every declared function is exported under libcd's own names; without a
    GPU the calls fail with -3 (no CPU path), a wrong cell type with libcd's -2."""
    import numpy as np
    from or_cdchomp_b200 import libcd
    src = open(os.path.join(ROOT, "include", "libcd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(cd_(?:grid|chomp)_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(libcd.EXPORTS)
    lib = libcd.load()
    for name in declared:
        assert getattr(lib, name) is not None
    # struct layout against the C compiler
    import subprocess
    import tempfile
    probe = '#include <stdio.h>\n#include "libcd_b200.h"\nint main(void){ printf("%zu\\n", sizeof(struct cd_grid)); return 0; }\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "p.c"), "w").write(probe)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "p.c"), "-o", os.path.join(d, "p")])
        assert int(subprocess.check_output([os.path.join(d, "p")])) == C.sizeof(libcd.CdGrid)
        probe2 = probe.replace("struct cd_grid", "struct cd_chomp")
        open(os.path.join(d, "q.c"), "w").write(probe2)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "q.c"), "-o", os.path.join(d, "q")])
        assert int(subprocess.check_output([os.path.join(d, "q")])) == C.sizeof(libcd.CdChomp)
    # cell type check comes first, as grid.c:648-649
    g = libcd.HostGrid(np.zeros((4, 4, 4)), [1, 1, 1], cell_size=1)
    out = C.POINTER(libcd.CdGrid)()
    assert lib.cd_grid_double_bin_sdf(C.byref(out), C.byref(g.c)) == -2
    g2 = libcd.HostGrid(np.zeros((4, 4)), [1, 1])
    assert lib.cd_grid_double_bin_sdf(C.byref(out), C.byref(g2.c)) == -2
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(libcd.LibcdError) as ei:
            libcd.bin_sdf(np.zeros((4, 4, 4)), [1, 1, 1])
        assert ei.value.code == -3 and "no CPU path" in str(ei.value)


def test_libcd_struct_layouts_match_reference_headers():
    """field offsets of struct cd_grid / struct cd_chomp in include/libcd_b200.h against the
    reference's own headers (only where the reference tree is mounted)."""
    ref = "/root/reference/src"
    if not os.path.isdir(os.path.join(ref, "libcd")):
        pytest.skip("reference tree not present")
    import subprocess
    import tempfile
    fields = ["n", "m", "lambda", "dt", "T", "ldt", "T_points", "G", "G_points", "AG", "AG_points", "D", "wds", "inits",
              "finals", "initsfinals", "A", "Ainv", "B", "trC", "jlimit_lower", "jlimit_upper", "Kvels", "Evels", "vels",
              "cost_nxn", "cost_mxn", "Gjlimit", "GjlimitAinv", "cptr", "cost_pre", "cost", "cost_extra", "use_momentum",
              "leapfrog_first", "cons", "cons_k", "cons_h", "cons_Jcol", "cons_JAJT", "cons_ipiv", "cons_delta",
              "ticks_vels", "ticks_callback_pre", "ticks_callbacks", "ticks_smoothgrad", "ticks_smoothcost"]
    gfields = ["n", "sizes", "ncells", "cell_size", "data", "lengths"]
    body = "".join('printf("%%zu\\n", offsetof(struct cd_chomp, %s));' % f for f in fields)
    body += "".join('printf("%%zu\\n", offsetof(struct cd_grid, %s));' % f for f in gfields)
    body += 'printf("%zu %zu\\n", sizeof(struct cd_chomp), sizeof(struct cd_grid));'
    head = "#include <stdio.h>\n#include <stdlib.h>\n#include <stddef.h>\n#include <time.h>\n"
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for name, inc, flags in (("ref", "#include <libcd/grid.h>\n#include <libcd/chomp.h>\n", ["-I", ref]),
                                 ("mine", '#include "libcd_b200.h"\n', ["-I", os.path.join(ROOT, "include")])):
            c = os.path.join(d, name + ".c")
            open(c, "w").write(head + inc + "int main(void){" + body + "return 0;}\n")
            subprocess.check_call(["gcc"] + flags + [c, "-o", os.path.join(d, name)])
            outs.append(subprocess.check_output([os.path.join(d, name)]))
    assert outs[0] == outs[1]


def test_kernel_sources_compile_under_nvrtc():
    """the persistent kernel must stay compilable without host headers: the engine compiles these
    very sources at run time (csrc/ocb_jit.cpp), generic and with the robot written out as
    constexpr tables (ocb_jit_robot.h, generated by the engine).  NVRTC needs no GPU."""
    try:
        from cuda.bindings import nvrtc
    except Exception:
        pytest.skip("cuda-python (nvrtc bindings) not importable")
    import ctypes as C
    from or_cdchomp_b200 import models
    lib = capi.load_library()
    lib.ocb_debug_jit_robot_header.restype = C.c_long
    csrc = os.path.join(ROOT, "or_cdchomp_b200", "csrc")
    src = open(os.path.join(csrc, "chomp_kernel.cu")).read()
    names = ["ocb_internal.h", "chomp_device.cuh", "chomp_jit_robot.cuh"]
    hdrs = {n: open(os.path.join(csrc, n)).read() for n in names}

    def robot_header(robot, params):
        buf = C.create_string_buffer(1 << 20)
        n = lib.ocb_debug_jit_robot_header(C.byref(robot.struct), C.byref(params), buf, len(buf))
        assert 0 < n < len(buf)
        return buf.value.decode()

    wam = robot_header(models.wam7_robot(), capi.default_params(n_points=100))
    assert "JR_NPA 96" in wam and "JR_HIT_WORDS 4" in wam   # 105 pairs - 9 on a shared link; + 15 x 1 inactive
    tree = robot_header(models.prismatic_test_robot(), capi.default_params(n_points=33, floating_base=1))
    # a robot with more sphere pairs than the hit set has bits keeps the table-driven kernel
    buf = C.create_string_buffer(16)
    assert lib.ocb_debug_jit_robot_header(C.byref(models.dense_sphere_arm(40).struct),
                                          C.byref(capi.default_params()), buf, len(buf)) == 0
    cases = [
        (dict(NT=128, MINBLOCKS=3, FLOAT=0, PP=100, NN=7, nsa=15, nsi=1, NAp=18, n_slots=0, ng=6, nj=7, nsdf=1,
              n_desc=21, use_momentum=0, use_hmc=0), ""),
        (dict(NT=64, MINBLOCKS=1, FLOAT=1, PP=0, NN=11, nsa=9, nsi=0, NAp=12, n_slots=2, ng=5, nj=6, nsdf=3,
              n_desc=14, use_momentum=1, use_hmc=1), ""),
        (dict(NT=128, MINBLOCKS=3, FLOAT=0, PP=100, NN=7, nsa=15, nsi=1, NAp=18, n_slots=0, ng=5, nj=7, nsdf=1,
              n_desc=22, use_momentum=0, use_hmc=0), wam),
        (dict(NT=64, MINBLOCKS=1, FLOAT=1, PP=33, NN=11, nsa=8, nsi=0, NAp=11, n_slots=1, ng=6, nj=6, nsdf=2,
              n_desc=14, use_momentum=1, use_hmc=1), tree),
    ]
    for defs, robot in cases:
        h = dict(hdrs)
        h["ocb_jit_robot.h"] = robot
        err, prog = nvrtc.nvrtcCreateProgram(src.encode(), b"chomp_kernel.cu", len(h),
                                             [v.encode() for v in h.values()], [k.encode() for k in h])
        assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS
        opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-DOCB_JIT=1", b"-default-device"]
        opts += [("-DOCB_JIT_%s=%d" % kv).encode() for kv in defs.items()]
        if robot:
            opts.append(b"-DOCB_JIT_ROBOT=1")
        err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
        _, n = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS, log.decode(errors="replace")[:2000]
        _, nb = nvrtc.nvrtcGetCUBINSize(prog)
        assert nb > 10000
        nvrtc.nvrtcDestroyProgram(prog)
    # every -D the kernel reads is one ocb_jit.cpp passes
    import re as _re
    used = set(_re.findall(r"OCB_JIT_([A-Za-z_]+)", src)) | set(_re.findall(r"DIM\(a, ([A-Za-z_]+)\)", src))
    jit_cpp = open(os.path.join(csrc, "ocb_jit.cpp")).read()
    passed = set(_re.findall(r'\{"([A-Za-z_]+)", ', jit_cpp)) | set(_re.findall(r"-DOCB_JIT_([A-Za-z_]+)=", jit_cpp))
    assert used - {"f", "FLAGS"} <= passed, used - passed


def test_metric_and_its_closed_forms():
    """host side of the smoothness metric (build_metric / metric_inverse, no device): against the reference's
    definition A = sum_d w_d / N_d K_d^T K_d, B = ... K_d^T E_d (chomp.c:239-340) built densely here, and the
    closed forms the kernels use for the default metric -- (A^-1)_ij = min(i,j)(m+1-max(i,j))/((m+1)c) and
    A^-1 (A T + B) = T - straight line"""
    import ctypes as C
    import numpy as np
    from or_cdchomp_b200 import capi
    lib = capi.load_library()

    def metric(P, D, free):
        m = P - 2 + free
        W = 2 * D + 1
        band, inv, bi, bf = np.zeros((m, W)), np.zeros((m, m)), np.zeros(m), np.zeros(m)
        flags, c = C.c_int(), C.c_double()
        assert lib.ocb_debug_metric(P, D, free, capi.dptr(band), capi.dptr(inv), capi.dptr(bi), capi.dptr(bf),
                                    C.byref(flags), C.byref(c)) == 0
        A = np.zeros((m, m))
        for i in range(m):
            for k in range(-D, D + 1):
                if 0 <= i + k < m:
                    A[i, i + k] = band[i, k + D]
        return A, inv, bi, bf, flags.value, c.value

    def dense_reference(P, D, free):
        """chomp.c:239-340 with wds = [0..0,1]: K_d by repeated differencing with boundary rows; inits[0] absent when free"""
        m, dt = P - 2 + free, 1.0 / (P - 1)
        K, Ei, Ef = np.eye(m), np.zeros(m), np.zeros(m)   # level -1: identity on the moving points
        for d in range(D):
            has_i = 0 if (d == 0 and free) else 1
            prev = K.shape[0]
            cur = prev - 1 + has_i + 1
            Dm = np.zeros((cur, prev))
            ci, cf = np.zeros(cur), np.zeros(cur)
            if has_i:
                Dm[0, 0] = 1.0 / dt
            for i in range(prev - 1):
                Dm[has_i + i, i], Dm[has_i + i, i + 1] = -1.0 / dt, 1.0 / dt
            Dm[cur - 1, prev - 1] = -1.0 / dt
            K, Ei, Ef = Dm @ K, Dm @ Ei, Dm @ Ef
            if d == 0:
                if has_i:
                    Ei[0] += -1.0 / dt
                Ef[cur - 1] += 1.0 / dt
        w = 1.0 / K.shape[0]
        return w * K.T @ K, w * K.T @ Ei, w * K.T @ Ef

    for P, D, free in ((100, 1, 0), (40, 1, 1), (50, 2, 0), (30, 3, 0), (5, 1, 0), (256, 1, 0)):
        A, inv, bi, bf, flags, c = metric(P, D, free)
        Ar, bir, bfr = dense_reference(P, D, free)
        m = A.shape[0]
        assert np.allclose(A, Ar, rtol=1e-13, atol=1e-9) and np.allclose(bi, bir, rtol=1e-13, atol=1e-9)
        assert np.allclose(bf, bfr, rtol=1e-13, atol=1e-9)
        assert np.allclose(inv @ A, np.eye(m), atol=1e-9)
        assert flags == (3 if (D == 1 and not free and m > 3) else 0), (P, D, free, flags)
        if flags == 3:
            i = np.arange(1, m + 1)
            closed = np.minimum.outer(i, i) * (m + 1 - np.maximum.outer(i, i)) / ((m + 1) * c)
            assert np.allclose(inv, closed, rtol=1e-11, atol=0)
            rng = np.random.default_rng(P)
            qs, qg, T = rng.normal(size=3), rng.normal(size=3), rng.normal(size=(m, 3))
            B = np.outer(bi, qs) + np.outer(bf, qg)
            line = qs[None] + (qg - qs)[None] * (i[:, None] / (m + 1))
            assert np.allclose(inv @ (A @ T + B), T - line, atol=1e-9)
    assert lib.ocb_debug_metric(2, 1, 0, None, None, None, None, None, None) != 0
