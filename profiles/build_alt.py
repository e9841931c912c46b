"""Experiments: build an alternative libdonk_b200 with extra nvcc flags for ONE translation unit (the others are taken
from build/obj).  python profiles/build_alt.py NAME abi_ilqr.cu -DDONK_PAIR_SPLIT ...  ->  profiles/_build/NAME.so"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g
name, unit, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
out = ROOT / "profiles" / "_build"; out.mkdir(exist_ok=True)
obj = out / f"{name}_{Path(unit).stem}.o"
subprocess.run(["nvcc"] + g.NVCC_FLAGS + flags + ["-c", "-o", str(obj), str(g.SRC / unit)], check=True, cwd=ROOT)
objs = [str(obj)] + [str(o) for o in sorted((ROOT / "build" / "obj").glob("*.o")) if o.stem != Path(unit).stem]
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out / f"{name}.so")] + objs, check=True)
print(out / f"{name}.so")
