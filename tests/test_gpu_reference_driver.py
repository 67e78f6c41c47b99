"""the reference's UNMODIFIED eval driver against the drop-in modules.

baseline/_ref/trainer_objective.py (copied byte for byte by baseline/install_ref.py) runs condGANTrainer.gen_exampleSRHL
with tgsr_b200/compat first on sys.path, on the shipped checkpoint and the real LR crop; the `*_SR.png` it writes must be
the uint8 image the reference's own modules produce for the same inputs (tests/golden/face_x8.pt holds that output):
identical up to 1 LSB with the fp32 kernels, within the 2e-2 bar (3 LSB) on the 16-bit tensor-core path."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(c, key):
    B = c['LR'].shape[0]
    pyr = [torch.zeros(B, 3, s, s) for s in (64, 128, 256)]
    imgs = [c['LR'].clone()] + pyr
    blur = [c['LR'].clone()] + [p.clone() for p in pyr]
    bic = [t.clone() for t in imgs]
    bicblur = [t.clone() for t in imgs]
    return (imgs, c['captions'].clone()[:, :, None], c['cap_lens'].clone(), torch.zeros(B, dtype=torch.int64), [key] * B,
            bic, blur, bicblur)


def _u8(x):
    return np.round(np.maximum(0, np.minimum(255, (x.numpy() + 1.0) * 127.5))).astype(np.uint8)


@pytest.mark.parametrize('precision,lsb', [('fp32', 1), ('mixed16', 3)])
def test_unmodified_driver_writes_the_reference_png(tmp_path, golden, precision, lsb):
    if not os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'trainer_objective.py')):
        pytest.skip('baseline/_ref not installed (python baseline/install_ref.py where /root/reference is mounted)')
    from PIL import Image
    w = golden('face_S8_weights.pt'); g = golden('face_x8.pt'); rnn = golden('rnn_encoder.pt')
    net_g = tmp_path / 'netG_epoch_7.pth'
    torch.save(w['netG'], net_g); torch.save(w['netGH'], tmp_path / 'netGH_epoch_7.pth')
    torch.save(rnn['sd'], tmp_path / 'text_encoder.pth')
    torch.save([_batch(g['b1'], 'b1'), _batch(g['b3'], 'b3')], tmp_path / 'batches.pt')
    (tmp_path / 'eval.yml').write_text(
        "CONFIG_NAME: 'attn2'\nDATASET_NAME: 'face'\nGPU_ID: 0\nCUDA: True\nTREE:\n  BRANCH_NUM: 4\n  BASE_SIZE: 32\n"
        "TRAIN:\n  FLAG: False\n  BATCH_SIZE: 1\n  NET_E: '%s'\n  NET_G: '%s'\n"
        "GAN:\n  DF_DIM: 64\n  GF_DIM: 32\n  Z_DIM: 100\n  R_NUM: 2\nTEXT:\n  EMBEDDING_DIM: 256\n  CAPTIONS_PER_IMAGE: 1\n  WORDS_NUM: 18\n"
        % (tmp_path / 'text_encoder.pth', net_g))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_run_reference_driver.py'), str(tmp_path), precision],
                       capture_output=True, text=True, timeout=600)
    assert 'DRIVER_DONE' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    out = tmp_path / 'netG_epoch_7' / 'sr_out'
    got1 = np.asarray(Image.open(out / 'b1_SR.png')).transpose(2, 0, 1)
    want1 = _u8(g['b1']['fine'][-1][0])
    d1 = np.abs(got1.astype(np.int32) - want1.astype(np.int32))
    print('%s: b1 SR png vs reference output: max %d LSB, %.4f%% pixels differ' % (precision, d1.max(), 100.0 * (d1 > 0).mean()))
    assert got1.shape == (3, 256, 256) and d1.max() <= lsb
    got3 = np.asarray(Image.open(out / 'b3_SR.png')).transpose(2, 0, 1)[:, ::4, ::4]       # image 0 of the ragged batch
    want3 = _u8(g['b3']['fine2_sub4'][0])
    assert np.abs(got3.astype(np.int32) - want3.astype(np.int32)).max() <= lsb
    assert (out / 'att' / 'b1.png').is_file()             # the unmodified attention visualisation ran as well
