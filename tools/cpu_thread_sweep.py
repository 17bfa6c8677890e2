import time, torch, os, sys
sys.path.insert(0, os.getcwd())
from oracle import unet3d_oracle as O
from videoswap_b200 import UNetConfig, seeded_state_dict, unet_param_shapes
sd = seeded_state_dict(unet_param_shapes(UNetConfig()), seed=0)
x = torch.randn(1, 4, 2, 64, 64); ehs = torch.randn(1, 16, 77, 768)
print("cpu_count", os.cpu_count(), flush=True)
for n in (16, 32, 64):
    torch.set_num_threads(n)
    with torch.no_grad():
        t0 = time.perf_counter(); O.unet_forward(sd, O.OracleConfig(), x, 981, ehs); t1 = time.perf_counter()
        O.unet_forward(sd, O.OracleConfig(), x, 981, ehs); t2 = time.perf_counter()
    print(n, "threads: first", round(t1 - t0, 2), "second", round(t2 - t1, 2), flush=True)
