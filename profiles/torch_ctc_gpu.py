import torch, time, sys
sys.path.insert(0,'/root/repo')
import bench
lp, il, tg, tl = bench.cfg2_inputs()
x = lp.cuda().transpose(0,1).contiguous().requires_grad_(True)   # (T,B,V) contiguous like most users
tg_d, il_d, tl_d = tg.cuda(), il.cuda(), tl.cuda()
def step():
    x.grad=None
    l = torch.nn.functional.ctc_loss(x, tg_d, il_d, tl_d, blank=1, reduction='mean', zero_infinity=True)
    l.backward()
for _ in range(5): step()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print("torch cuda ctc_loss fwd+bwd ms/step:", e0.elapsed_time(e1)/20)
# view like the reference: (B,T,V) contiguous transposed
y = lp.cuda().requires_grad_(True)
def step2():
    y.grad=None
    l = torch.nn.functional.ctc_loss(y.transpose(0,1), tg_d, il, tl_d, blank=1, reduction='mean', zero_infinity=True)
    l.backward()
for _ in range(5): step2()
torch.cuda.synchronize(); e0.record()
for _ in range(20): step2()
e1.record(); torch.cuda.synchronize()
print("torch cuda ctc_loss (reference call shape, lengths on cpu) ms/step:", e0.elapsed_time(e1)/20)
