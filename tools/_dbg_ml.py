import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import pram_oracle as O, ref_loader as RL
from pram_b200 import ops
from pram_b200.nets.gml import GML
dev = torch.device('cuda')
g = torch.Generator().manual_seed(0)
for (b, m, n) in [(2, 4096, 1024), (20, 4096, 1024), (2, 1024, 4096)]:
    dist = torch.randn(b, m, n, generator=g) * 2
    for i in range(min(m, n) // 2):
        dist[:, i, (i * 3) % n] += 15
    bs = torch.tensor(0.7)
    m0, m1, s0, s1 = ops.sinkhorn_match(dist.to(dev), bs.to(dev), 20, 0.2)
    P = O.sinkhorn_with_dustbin(dist[:1], bs, 20)
    i0, i1, t0, t1 = O.compute_matches(P, 0.2)
    print('sinkhorn', b, m, n, 'matched ours', int((m0[0] > -1).sum()), 'oracle', int((i0[0] > -1).sum()),
          'score diff', float((s0[0].cpu() - t0[0]).abs().max()), 'last pair matched', int((m0[-1] > -1).sum()))
sd = RL.load_gml_state()
net = GML({}); net.load_state_dict(sd, strict=True); net = net.to(dev)
m, n = 4096, 1024
d0 = torch.nn.functional.normalize(torch.randn(1, m, 128, generator=g), dim=-1)
k0 = torch.rand(1, m, 2, generator=g) * torch.tensor([1600., 1200.])
perm = torch.randperm(m, generator=g)[:n]
data = {'descriptors0': d0, 'descriptors1': d0[:, perm], 'keypoints0': k0, 'keypoints1': k0[:, perm],
        'image_shape0': (1, 3, 1600, 1200), 'image_shape1': (1, 3, 1600, 1200)}
for b in (1, 4):
    dd = {k: (v.repeat(b, 1, 1).to(dev) if torch.is_tensor(v) else v) for k, v in data.items()}
    out = net(dd)
    print('gml b', b, 'matched', [(int((out['matches0'][i] > -1).sum())) for i in range(b)])
ref = O.gml_forward(sd, data)
print('oracle matched', int((ref['matches0'] > -1).sum()), 'agree', float((out['matches0'][0].cpu() == ref['matches0'][0]).float().mean()))
