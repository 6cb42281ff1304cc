"""Compile the run-time specialised kernel with NVRTC itself (cuda-python bindings), without a GPU:
catches what nvcc accepts and NVRTC does not.  Usage: dev_nvrtc_check.py [generic]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuda.bindings import nvrtc
from or_cdchomp_b200 import capi, models
CSRC = os.path.join(ROOT, "or_cdchomp_b200", "csrc")
lib = capi.load_library()
lib.ocb_debug_jit_robot_header.restype = C.c_long


def check(robot, params, defs, generic=False):
    buf = C.create_string_buffer(1 << 20)
    n = lib.ocb_debug_jit_robot_header(C.byref(robot.struct), C.byref(params), buf, len(buf))
    hdr = buf.value if (n > 0 and not generic) else b""
    names = [b"ocb_internal.h", b"chomp_device.cuh", b"chomp_jit_robot.cuh", b"ocb_jit_robot.h"]
    srcs = [open(os.path.join(CSRC, n.decode()), "rb").read() for n in names[:3]] + [hdr]
    err, prog = nvrtc.nvrtcCreateProgram(open(os.path.join(CSRC, "chomp_kernel.cu"), "rb").read(), b"chomp_kernel.cu",
                                         4, srcs, names)
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"-DOCB_JIT=1", b"-default-device"] + [("-DOCB_JIT_%s=%d" % kv).encode() for kv in defs.items()]
    if hdr:
        opts.append(b"-DOCB_JIT_ROBOT=1")
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    _, n = nvrtc.nvrtcGetProgramLogSize(prog)
    log = b" " * n
    nvrtc.nvrtcGetProgramLog(prog, log)
    print("robot mode" if hdr else "generic", "rc", err, log.decode()[:3000])
    return int(err)


if __name__ == "__main__":
    generic = "generic" in sys.argv
    rc = check(models.wam7_robot(), capi.default_params(n_points=100), dict(
        NT=128, MINBLOCKS=3, FLOAT=0, PP=100, NN=7, nsa=15, nsi=1, NAp=18, n_slots=0, ng=5, nj=7, nsdf=1, n_desc=22,
        use_momentum=0, use_hmc=0), generic)
    # tree robot with prismatic / mimic joints, momentum + hmc, floating base
    rb = models.prismatic_test_robot()
    rc |= check(rb, capi.default_params(n_points=33, floating_base=1), dict(
        NT=64, MINBLOCKS=1, FLOAT=1, PP=33, NN=11, nsa=8, nsi=0, NAp=11, n_slots=1, ng=6, nj=6, nsdf=2, n_desc=14,
        use_momentum=1, use_hmc=1), generic)
    sys.exit(rc)
