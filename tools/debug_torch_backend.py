import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import kbnet_b200, kbnet_b200.model as M
from oracle import kbnet_oracle as K
torch.set_grad_enabled(False)
g = torch.load('tests/golden/model_small.pt', weights_only=False)
dev = torch.device('cuda:0')
for backend in ('kbnet', 'torch'):
    M.CONV_BACKEND = backend
    m = kbnet_b200.KBNetModel(3, 2, g['min_pool_sizes'], g['max_pool_sizes'], 3, 8, g['f_img'], g['f_dep'], [0,1,2,3], g['f_dec'], device=dev)
    m.load_state_dicts(g['s2d_state_dict'], g['enc_state_dict'], g['dec_state_dict'])
    enc = m.encoder
    img, dep = g['image'].to(dev), g['s2d_out'].to(dev)
    w = g['enc_state_dict']
    c0i = M._conv([img], enc.conv0_image.weight, 1, 0.2)
    r0i = K.conv2d(g['image'], w['conv0_image.conv.weight'])
    print(backend, 'conv0_image err', float((c0i.cpu()-r0i).abs().max()/r0i.abs().max()))
    latent, skips = enc(img, dep, g['intrinsics'].to(dev))
    for i,(a,b) in enumerate(zip(skips, g['skips'])):
        print(backend, 'skip', i, float((a.cpu()-b).abs().max()/b.abs().max()))
    print(backend, 'latent', float((latent.cpu()-g['latent']).abs().max()/g['latent'].abs().max()), latent.shape, latent.stride())
    f5 = enc.conv5_image.weight.shape[0]
    print(backend, 'latent img part', float((latent[:, :f5].cpu()-g['latent'][:, :f5]).abs().max()), 'dep part', float((latent[:, f5:].cpu()-g['latent'][:, f5:]).abs().max()))
