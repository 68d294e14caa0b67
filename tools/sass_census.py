#!/usr/bin/env python
"""Mnemonic census + a TMA excerpt per kernel from `cuobjdump -sass` of the built library
(writes profiles/r02_sass_excerpts.md; no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ouster-sdk_b200", "lib", "libouster_b200.so")
WANT = [("decode_pipe_kernelIfLi864", "decode_pipe_kernel<float,864> (K2, 72-register build)"),
        ("decode_pipe_kernelIfLi1024", "decode_pipe_kernel<float,1024> (K2, 64-register build: 2-3 CTAs/SM, helper warps)"),
        ("cloud_tma_kernelIfLi2ELb0ELb0", "cloud_tma_kernel<float,2,false,false> (K1)"),
        ("cloud_tma_kernelIfLi2ELb1ELb0", "cloud_tma_kernel<float,2,true,false> (K1, poses fused)"),
        ("cloud_tma_kernelIfLi2ELb0ELb1", "cloud_tma_kernel<float,2,false,true> (K1, LUT-free)"),
        ("k3_fused_kernelIf", "k3_fused_kernel<float> (K3)"),
        ("encode_kernel", "encode_kernel (K4)"),
        ("normals_kernelIf", "normals_kernel<float>")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = {}, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        funcs[cur].append(ln.split("/*", 2)[0] + "/*" + ln.split("/*", 2)[1] if False else ln)
out = ["# SASS evidence, round 2 (`cuobjdump -sass ouster-sdk_b200/lib/libouster_b200.so`, sm_100a; `tools/sass_census.py`)", "",
       "Mnemonic census per kernel (TMA: `UBLKCP` = 1-D bulk copy, `UTMALDG` = tensor-map copy; `SYNCS` = mbarrier ops; "
       "no tensor-core instructions (`HMMA`/`UTCMMA`/`QGMMA` absent): the path has no contraction).", ""]
KEEP = ("UBLKCP", "UTMALDG", "UBLKPF", "SYNCS", "LDS", "STS", "STG", "LDG", "ATOM", "RED", "SHFL", "ST.E", "LD.E", "LDL", "STL", "MMA")
for key, title in WANT:
    name = next((f for f in funcs if key in f), None)
    if not name:
        continue
    ins = funcs[name]
    ops = collections.Counter()
    for ln in ins:
        body = re.sub(r"/\*[0-9a-f]+\*/", "", ln).strip().rstrip(";").strip()
        body = re.sub(r"^@!?U?P\d+\s+", "", body)
        op = body.split(" ")[0]
        if any(k in op for k in KEEP):
            ops[op] += 1
    out += [f"## {title}", "", f"{len(ins)} instructions; " + ", ".join(f"`{k}` x{v}" for k, v in ops.most_common(18)), ""]
    idx = next((i for i, ln in enumerate(ins) if "UTMALDG" in ln), None)
    if idx is None:
        idx = next((i for i, ln in enumerate(ins) if "UBLKCP" in ln), None)
    if idx is not None:
        out += ["```"] + [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in ins[max(0, idx - 5):idx + 6]] + ["```", ""]
open(os.path.join(ROOT, "profiles", "r02_sass_excerpts.md"), "w").write("\n".join(out) + "\n")
print("\n".join(l for l in out if l.startswith("##") or "instructions;" in l))
