import torch, numpy as np, os, subprocess
print(torch.__version__, torch.backends.cpu.get_cpu_capability(), os.cpu_count(), torch.get_num_threads())
print(subprocess.run("lscpu | grep -E 'Model name|Flags' | cut -c1-400", shell=True, capture_output=True, text=True).stdout)
x = torch.tensor([-2.0032], dtype=torch.float32)
print("exp1", repr(torch.exp(x).item()))
xs = torch.full((64,), -2.0032, dtype=torch.float32)
print("exp64", repr(torch.exp(xs)[0].item()))
y = torch.arange(0,512); xx=torch.arange(0,512)
yy,xg = torch.meshgrid(y,xx,indexing="ij")
q = -((xg-100)**2+(yy-100)**2)/(2*50**2)
print("q", q.dtype, repr(float(q[96,0])), "exp(q)", repr(float(torch.exp(q)[96,0])))
for t in (1, 4, 16):
    torch.set_num_threads(t)
    print("threads", t, repr(float(torch.exp(q)[96,0])), repr(float(torch.exp(q.contiguous().clone())[96,0])))
