"""Stand-in for skimage.metrics (0.19.x) -- forwards to the oracle restatement."""
from oracle.metrics import ssim_skimage


def structural_similarity(im1, im2, win_size=None, **kw):
    assert not kw, kw
    return ssim_skimage(im1, im2, win_size=7 if win_size is None else win_size)
