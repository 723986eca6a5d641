"""ORACLE (test infrastructure): numpy restatement of OpenCV's 8-bit INTER_LINEAR resize, the arithmetic behind
`cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)` that the reference calls in agent.py:100-103,
inverse_dynamics_model.py:54-59 and data_loader.py:113-120.  OpenCV (4.x, modules/imgproc/src/resize.cpp) is a third-party
dependency of the reference (requirements.txt: opencv-python); its published algorithm for uint8:
  fx = (dx + 0.5) * scale - 0.5; sx = floor(fx); fx -= sx; clamp at the borders;
  weights = round(w * 2^11) as int16; horizontal pass in int32; vertical pass
  dst = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
Pinned bit-exact against cv2 itself in tests/test_agent.py wherever cv2 is importable -- for DOWNSCALING (all the reference
does: 640x360 / 1280x720 -> 128x128); when upscaling OpenCV takes another code path and differs by +-1 from this formula."""
import numpy as np


def tables(dst, src):
    scale = src / dst
    idx = np.zeros(dst, np.int64)
    w = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        f = (d + 0.5) * scale - 0.5
        s = int(np.floor(f))
        f -= s
        if s < 0:
            s, f = 0, 0.0
        if s >= src - 1:
            s, f = src - 1, 0.0
        idx[d] = s
        w[d] = (int(np.rint((1.0 - f) * 2048)), int(np.rint(f * 2048)))
    return idx, w


def resize_linear_u8(img, dw, dh):
    """img uint8 [H, W, C] -> uint8 [dh, dw, C]."""
    H, W, _ = img.shape
    xi, xa = tables(dw, W)
    yi, ya = tables(dh, H)
    src = img.astype(np.int64)
    x1 = np.minimum(xi + 1, W - 1)
    hor = src[:, xi, :] * xa[None, :, 0, None] + src[:, x1, :] * xa[None, :, 1, None]
    y1 = np.minimum(yi + 1, H - 1)
    out = ((ya[:, 0, None, None] * (hor[yi] >> 4)) >> 16) + ((ya[:, 1, None, None] * (hor[y1] >> 4)) >> 16)
    return np.clip((out + 2) >> 2, 0, 255).astype(np.uint8)


def composite_cursor(frame, cursor, alpha, x, y):
    """data_loader.py:34-45 (composite_images_with_alpha): frame uint8 [H, W, 3] modified in place; cursor uint8 [ch, cw, 3];
    alpha float64 [ch, cw, 1]; float64 arithmetic, astype(uint8) truncation, overlay clipped at the right / bottom border."""
    ch = max(0, min(frame.shape[0] - y, cursor.shape[0]))
    cw = max(0, min(frame.shape[1] - x, cursor.shape[1]))
    if ch == 0 or cw == 0:
        return frame
    a = alpha[:ch, :cw]
    frame[y:y + ch, x:x + cw, :] = (frame[y:y + ch, x:x + cw, :] * (1 - a) + cursor[:ch, :cw, :] * a).astype(np.uint8)
    return frame


def ingest(frame_bgr, size, cursor=None, alpha=None, xy=None):
    """data_loader.py:108-118: [cursor overlay] -> BGR2RGB -> resize (INTER_LINEAR) -> uint8 [size[1], size[0], 3]."""
    f = frame_bgr.copy()
    if cursor is not None and xy is not None and xy[0] >= 0:
        composite_cursor(f, cursor, alpha, int(xy[0]), int(xy[1]))
    return resize_linear_u8(f[:, :, ::-1], size[0], size[1])
