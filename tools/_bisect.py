import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from tests import unet_checks as U, kernel_checks as KC
from videoswap_b200 import ops
def run(tag, **opts):
    for k, v in opts.items(): ops.set_option(k, v)
    try:
        r = U.unet_vs_reference_golden("full_arch_c1")
    except Exception as e:
        r = {"error": repr(e)[:200]}
    for k in opts: ops.set_option(k, {"attn_persist":1,"subpixel":1,"ln_fuse":1,"attn_epiwg":1,"attn_poly":1,"gemm_pair":1}[k])
    print(tag, opts, json.dumps(r), flush=True)
run("default")
run("no_persist", attn_persist=0)
run("no_subpixel", subpixel=0)
run("no_lnfuse", ln_fuse=0)
run("no_epiwg", attn_epiwg=0)
for B in (1, 2):
    for name, fn in (("self_d40_n4096", lambda: KC.check_self_attention(B=B, N=4096, C=320, seed=124)),
                     ("self_d80_n1024", lambda: KC.check_self_attention(B=B, N=1024, C=640, seed=126)),
                     ("cross_d40", lambda: KC.check_cross_attention(B=B, Fr=1, N=4096, C=320, seed=131)),
                     ("cross_d80", lambda: KC.check_cross_attention(B=B, Fr=1, N=1024, C=640, seed=131))):
        print("B", B, name, fn(), flush=True)
r = U.unet_vs_oracle(B=1, Fr=1, hw=64, edlora=True, taps=True)
print("oracle taps", json.dumps({k: round(v["psnr"], 1) for k, v in r["taps"].items()}), r["psnr"])
