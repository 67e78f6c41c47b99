import sys, torch
sys.path.insert(0, '.')
import tgsr_b200
from oracle import tgsr_oracle as O
from tgsr_b200.pipeline import SRPipeline
tgsr_b200.set_eval_defaults()
def check(variant, B, size, prec, ragged):
    tgsr_b200.set_precision(prec)
    from tgsr_b200.config import cfg
    cfg.TREE.BRANCH_NUM = 4 if variant == 'x8' else 5
    pipe = SRPipeline(41, device='cuda:0', variant=variant, low='lr', seed=100, randomize_bn=True)
    LR = O.synth_lr(B, size, seed=7); caps, lens = O.synth_captions(B, 41, seed=8, ragged=ragged)
    fine, imgsL, atts = pipe.forward_device(LR.cuda(), caps.cuda(), lens.cuda())
    torch.cuda.synchronize()
    sdE = {k: v.cpu() for k, v in pipe.text_encoder.state_dict().items()}
    sdL = {k: v.cpu() for k, v in pipe.netGL.state_dict().items()}; sdH = {k: v.cpu() for k, v in pipe.netGH.state_dict().items()}
    words, sent = O.rnn_encoder(sdE, caps, lens); mask = (caps == 0)[:, :words.shape[2]]
    r_imgs, r_atts, _, _, _ = O.g_sr_net_low(sdL, LR, sent, words, mask, variant=variant)
    r_fine = O.netg_high(sdH, LR, r_imgs, LR, low='lr', variant=variant)
    err = max((a.cpu() - b).abs().max().item() for a, b in zip(fine, r_fine))
    print('%s B=%d size=%d %s ragged=%s: max-abs %.3e, n_out %d' % (variant, B, size, prec, ragged, err, len(fine)))
    assert err < (1e-3 if prec == 'fp32' else 3e-2), err
for args in [('x8', 5, 32, 'mixed16', True), ('x8', 33, 32, 'mixed16', False), ('x8', 1, 32, 'mixed16', False), ('x8', 3, 32, 'fp32', True),
             ('x16', 2, 16, 'mixed16', True), ('x16', 2, 16, 'fp32', False), ('x8', 2, 48, 'mixed16', False)]:
    try:
        check(*args)
    except Exception as e:
        print('FAIL', args, repr(e)[:300])
