"""GPU geometric transforms (ramdsir_b200.transforms) against Pillow at run time and against goldens recorded from the
reference's own transform classes (oracle/gen_transform_golden.py imports dataset/transform.py unmodified).
Byte work: bit-exact.  RNG consumption: the next random.random() after the transform equals the reference's."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch
from PIL import Image

import fixture_data as FX

pytestmark = pytest.mark.gpu


def _sample(seed, h, w):
    """Same construction as oracle/gen_transform_golden.py::make_sample."""
    rs = np.random.RandomState(1000 + seed)
    img = (FX._blobs(rs, h, w, 3) * 255).round().astype(np.uint8).transpose(1, 2, 0)
    yy, xx = np.mgrid[:h, :w]
    r = np.hypot(yy - h / 2, xx - w / 2)
    mask = np.full((h, w), 255, np.uint8)
    mask[r < min(h, w) * 0.30] = 128
    mask[r < min(h, w) * 0.15] = 0
    return np.ascontiguousarray(img), mask


@pytest.mark.parametrize("shape", [(300, 280, 256, 256), (800, 801, 256, 256), (97, 131, 200, 150), (256, 256, 384, 300),
                                   (64, 200, 64, 77), (123, 45, 17, 45), (5, 7, 40, 33), (256, 256, 256, 256),
                                   (1634, 1634, 256, 256)])
def test_resize_matches_pillow(shape):
    from ramdsir_b200 import transforms as T
    hin, win, hout, wout = shape
    rs = np.random.RandomState(hin * 7 + win)
    img = rs.randint(0, 256, (hin, win, 3)).astype(np.uint8)
    mask = rs.randint(0, 256, (hin, win)).astype(np.uint8)
    got = T.resize_u8(torch.from_numpy(img).cuda(), (wout, hout), T.BILINEAR).cpu().numpy()
    np.testing.assert_array_equal(got, np.array(Image.fromarray(img).resize((wout, hout), Image.BILINEAR)))
    got_m = T.resize_u8(torch.from_numpy(mask).cuda(), (wout, hout), T.NEAREST).cpu().numpy()
    np.testing.assert_array_equal(got_m, np.array(Image.fromarray(mask).resize((wout, hout), Image.NEAREST)))
    # single-channel bilinear (mode "L") as well
    got_l = T.resize_u8(torch.from_numpy(mask).cuda(), (wout, hout), T.BILINEAR).cpu().numpy()
    np.testing.assert_array_equal(got_l, np.array(Image.fromarray(mask).resize((wout, hout), Image.BILINEAR)))


def test_transforms_match_the_reference_classes(golden_dir):
    from ramdsir_b200 import transforms as T
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    sha = lambda a: np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)
    for seed, h, w in g["cases"]:
        img, mask = _sample(int(seed), int(h), int(w))
        dimg, dmask = torch.from_numpy(img).cuda(), torch.from_numpy(mask).cuda()
        for name in ("resize", "scalecrop", "both"):
            random.seed(500 + int(seed))
            s = {"img": dimg, "mask": dmask}
            if name == "resize":
                s = T.Resize((256, 256))(s)
            elif name == "scalecrop":
                s = T.RandomScaleCrop((256, 256))(s)
            else:
                s = T.RandomScaleCrop((256, 256))(T.Resize((256, 256))(s))         # train.py:541
            oi, om = s["img"].cpu().numpy(), s["mask"].cpu().numpy()
            key = "%s_%d" % (name, seed)
            assert tuple(oi.shape[:2]) == tuple(g[key + "_shape"]), key
            np.testing.assert_array_equal(oi[::4, ::4], g[key + "_img_sub"], err_msg=key)
            np.testing.assert_array_equal(om[::4, ::4], g[key + "_mask_sub"], err_msg=key)
            np.testing.assert_array_equal(sha(oi), g[key + "_img_sha"], err_msg=key)
            np.testing.assert_array_equal(sha(om), g[key + "_mask_sha"], err_msg=key)
            assert random.random() == float(g[key + "_after"][0]), key          # same draws consumed, same order


def test_device_side_loader_chain_equals_host_chain():
    """decode -> Resize -> RandomScaleCrop -> CHW planes -> RAM, all on the device from the uint8 HWC image, equals the
    reference's host chain (PIL transforms, np.array, transpose) followed by the same RAM call."""
    import ramdsir_b200 as RR
    from ramdsir_b200 import transforms as T
    img, mask = _sample(11, 333, 301)
    par, _ = _sample(12, 640, 600)
    random.seed(77)
    s = T.RandomScaleCrop((256, 256))(T.Resize((256, 256))({"img": torch.from_numpy(img).cuda(), "mask": torch.from_numpy(mask).cuda()}))
    src_chw = T.to_chw(s["img"])[None]
    partner = T.to_chw(T.resize_u8(torch.from_numpy(par).cuda(), (256, 256), T.BILINEAR))[None]       # fundus.py:209
    random.seed(77)
    pil = {"img": Image.fromarray(img).resize((256, 256), Image.BILINEAR), "mask": Image.fromarray(mask).resize((256, 256), Image.NEAREST)}
    if random.random() > 0.5:
        w = int(random.uniform(1, 1.5) * 256)
        h = int(random.uniform(1, 1.5) * 256)
        pil = {"img": pil["img"].resize((w, h), Image.BILINEAR), "mask": pil["mask"].resize((w, h), Image.NEAREST)}
    w, h = pil["img"].size
    x, y = random.randint(0, w - 256), random.randint(0, h - 256)
    host_img = np.array(pil["img"].crop((x, y, x + 256, y + 256))).transpose(2, 0, 1)
    host_par = np.array(Image.fromarray(par).resize((256, 256), Image.BILINEAR)).transpose(2, 0, 1)
    assert torch.equal(src_chw[0].cpu(), torch.from_numpy(np.ascontiguousarray(host_img)))
    assert torch.equal(partner[0].cpu(), torch.from_numpy(np.ascontiguousarray(host_par)))
    bank = RR.build_amp_bank(partner)
    got = RR.ram_mix_batch(src_chw, bank, [0], [4], clip=(0.0, 255.0), scale=1 / 127.5, shift=-1.0)
    want = RR.ram_mix_batch(torch.from_numpy(np.ascontiguousarray(host_img)).cuda()[None], RR.build_amp_bank(torch.from_numpy(np.ascontiguousarray(host_par)).cuda()[None]),
                            [0], [4], clip=(0.0, 255.0), scale=1 / 127.5, shift=-1.0)
    assert torch.equal(got, want)
