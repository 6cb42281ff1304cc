"""Compile the run-time specialised kernel offline (nvcc instead of NVRTC, same source and defines)
for the WAM7 bench configuration: registers, SASS size, instruction mix.  No GPU needed."""
import ctypes as C, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from or_cdchomp_b200 import capi, models
lib = capi.load_library()
robot = models.wam7_robot()
params = capi.default_params(n_points=100, lambda_=100.0, obs_factor=500.0)
lib.ocb_debug_jit_robot_header.restype = C.c_long
buf = C.create_string_buffer(1 << 20)
n = lib.ocb_debug_jit_robot_header(C.byref(robot.struct), C.byref(params), buf, len(buf))
out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/jr"
os.makedirs(out, exist_ok=True)
robot_mode = n > 0 and "--generic" not in sys.argv
open(os.path.join(out, "ocb_jit_robot.h"), "w").write(buf.value.decode())
defs = dict(NT=128, MINBLOCKS=3, FLOAT=0, PP=100, NN=7, nsa=15, nsi=1, NAp=18, n_slots=0, ng=5, nj=7, nsdf=1,
            n_desc=22, use_momentum=0, use_hmc=0)
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-DOCB_JIT=1",
       "-D__CUDACC_RTC_SIM__", "-Xptxas", "-v", "-cubin", "-o", os.path.join(out, "k.cubin"),
       "-I", out, "-I", os.path.join(ROOT, "or_cdchomp_b200", "csrc")]
cmd += ["-DOCB_JIT_%s=%d" % kv for kv in defs.items()]
if robot_mode:
    cmd.append("-DOCB_JIT_ROBOT=1")
cmd += os.environ.get("OCB_JIT_FLAGS", "").split()
cmd.append(os.path.join(ROOT, "or_cdchomp_b200", "csrc", "chomp_kernel.cu"))
r = subprocess.run(cmd, capture_output=True, text=True)
import re
print('\n'.join(l for l in (r.stdout + r.stderr).splitlines() if re.search(r'error|warning|chomp_iterate_jit|Used 1|spill', l))[-3000:])
if r.returncode == 0:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(out, "k.cubin")], capture_output=True, text=True).stdout
    per = collections.defaultdict(collections.Counter)
    fn = None
    for ln in sass.splitlines():
        ln = ln.strip()
        if ln.startswith("Function :"):
            fn = ln.split(":", 1)[1].strip()
        if ln.startswith("/*") and ";" in ln and fn:
            body = ln.split("*/", 1)[1].strip()
            toks = body.split()
            if toks and toks[0].startswith("@"):
                toks = toks[1:]
            if toks:
                per[fn][toks[0].split(".")[0]] += 1
    for fn, ops in per.items():
        if "chomp_iterate" not in fn and "sdf_axis" not in fn:
            continue
        print(fn[-40:], "SASS instructions:", sum(ops.values()), "robot mode" if robot_mode else "generic")
        print("   " + " ".join("%s=%d" % kv for kv in ops.most_common(22)))
