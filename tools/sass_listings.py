"""Writes the SASS listing of every hot kernel of the built library to profiles/sass/<name>.sass (one file per kernel,
demangled name + source-less instruction listing + opcode histogram header).  usage: python tools/sass_listings.py"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "arraymancer_b200", "libarraymancer_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WANT = [  # (file stem, regex on the demangled name; first match wins)
    ("gemm_tf32x3_kernel_2cta", r"gemm_tf32x3_kernel<2>"),
    ("split_pack_vec_kernel", r"split_pack_vec_kernel"),
    ("contract_dmma_tma_kernel_kmaj_kmaj", r"contract_dmma_tma_kernel<true, true"),
    ("contract_dmma_tma_kernel_mnmaj_mnmaj", r"contract_dmma_tma_kernel<false, false"),
    ("contract_simt_i32_128x128", r"contract_simt_kernel<int, am::SimtCfg<int, 128, 128, 16, 4, 8>, am::StridedLoader<int>, am::StridedLoader<int>, am::StridedEpilogue<int>"),
    ("contract_simt_i64_128x128", r"contract_simt_kernel<long, am::SimtCfg<long, 128, 128, 8, 4, 8>, am::StridedLoader<long>, am::StridedLoader<long>, am::StridedEpilogue<long>"),
    ("skinny_kn_kernel_f32_16", r"skinny_kn_kernel<float, 16>"),
    ("skinny_nk_kernel_f32_16", r"skinny_nk_kernel<float, 16>"),
    ("conv_tc_kernel_unchecked", r"conv_tc_kernel<false>"),
    ("conv_dgrad_tc_kernel", r"conv_dgrad_tc_kernel"),
    ("conv_wgrad_tc_kernel", r"conv_wgrad_tc_kernel"),
    ("conv_c1_forward_kernel_5x5", r"conv_c1_forward_kernel<5, 5, false>"),
    ("conv_c1_backward_kernel_5x5", r"conv_c1_backward_kernel<5, 5>"),
    ("umma_pattern_kernel_all", r"umma_pattern_kernel<127>"),
]
funcs, cur = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    if cur:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m: funcs[cur].append((m.group(1), m.group(2).strip()))
names = list(funcs)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
os.makedirs(os.path.join(ROOT, "profiles", "sass"), exist_ok=True)
for stem, rx in WANT:
    hit = next(((n, d) for n, d in zip(names, dem) if re.search(rx, d)), None)
    if not hit:
        print("missing", stem); continue
    n, d = hit
    ins = funcs[n]
    h = collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]).split(".")[0] for _, i in ins)
    with open(os.path.join(ROOT, "profiles", "sass", stem + ".sass"), "w") as f:
        f.write(f"// {d}\n// {n}\n// {len(ins)} instructions; top opcodes: {dict(h.most_common(16))}\n")
        for a, i in ins: f.write(f"/*{a}*/ {i}\n")
    print(stem, len(ins))
