import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import problem, compare
from general_motion_retargeting_b200.synthetic import make_clips
from general_motion_retargeting_b200 import GeneralMotionRetargeting
from oracle import native
np.set_printoptions(linewidth=200, precision=2)
for src,robot in (('smplx','engineai_pm01'),):
    m, tt, _ = problem(src, robot)
    clips = make_clips(m, tt, range(16), T=40, src_human=src)
    g = GeneralMotionRetargeting(src, robot)
    pos, quat, h = torch.from_numpy(clips.pos).cuda(), torch.from_numpy(clips.quat).cuda(), torch.from_numpy(clips.heights).cuda()
    for prec in ('f32','f64'):
        q, it, err = g.retarget_batch(pos, quat, h, return_info=True, precision=prec)
        q = q.double().cpu().numpy(); it = it.cpu().numpy()
        for flags,name in ((0,'literal'),(0x10000,'stable')):
            q_ref, it_ref, _ = native.retarget_batch(m, tt, clips.pos, clips.quat, clips.ratio(tt), flags=flags)
            print(prec, name, compare(q, it, q_ref, it_ref))
        d = np.abs(q-q_ref).max(-1)
        c = int(d.max(1).argmax()); print('  worst clip', c, 'per frame', d[c]); print('  iters gpu', it[c].tolist()); print('  iters ref', it_ref[c].tolist())
