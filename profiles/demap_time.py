import sys, torch
sys.path.insert(0, "/root/repo/esn-ofdm-mimo_b200")
import esn_b200
from esn_b200 import ofdm
esn_b200.load()
B, N, Nt = 9472, 512, 4
y = torch.randn(B, N, 2*Nt, device="cuda")
tx = torch.randint(0, 16, (B, N, Nt), device="cuda", dtype=torch.uint8)
counts = torch.zeros(2, dtype=torch.int64, device="cuda")
big = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
ts=[]
for i in range(6):
    big.zero_()  # flush L2
    a,b_=ev(),ev(); a.record()
    ofdm.unpack_fft_demap(y, N, Nt, 1e-4, 4, tx_idx=tx, want_xhat=False, counts=counts)
    b_.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b_))
byt = y.numel()*4 + 2*tx.numel()
print("unpack_fft_demap: %.1f us  -> %.0f GB/s algorithmic (%.1f MB)" % (min(ts)*1e3, byt/min(ts)/1e6, byt/1e6), ts)
