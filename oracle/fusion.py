"""ORACLE (test infrastructure, never imported by the product): NumPy restatement of the Gipuma fusibile consensus kernel,
jdacs/fusion/fusibile/fusibile.cu:138-277 with helpers :46-133 (third-party code vendored in the reference tree).

PARITY UNPINNED: the fusibile binary needs cmake + OpenCV C++ and a GPU, neither of which this container offers, and the
reference holds no golden output for it.  The restatement follows the source line by line; the texture fetches
(cudaFilterModeLinear, un-normalised coordinates, main.cpp:489-493) follow the CUDA programming guide's definition of linear
filtering (texel-centre bilinear, clamped addresses, 8 fractional weight bits)."""
import numpy as np


def tex_linear(img, x, y):
    h, w = img.shape[:2]
    fx, fy = np.floor(x), np.floor(y)
    ax = np.floor((np.float32(x) - np.float32(fx)) * np.float32(256) + np.float32(0.5)) / np.float32(256)
    ay = np.floor((np.float32(y) - np.float32(fy)) * np.float32(256) + np.float32(0.5)) / np.float32(256)
    x0, x1 = int(np.clip(fx, 0, w - 1)), int(np.clip(fx + 1, 0, w - 1))
    y0, y1 = int(np.clip(fy, 0, h - 1)), int(np.clip(fy + 1, 0, h - 1))
    one = np.float32(1)
    return ((one - ax) * (one - ay) * img[y0, x0] + ax * (one - ay) * img[y0, x1] + (one - ax) * ay * img[y1, x0] + ax * ay * img[y1, x1]).astype(np.float32)


def point_of(cam, px, py, depth):
    m_inv, col = cam[12:21].reshape(3, 3), cam[21:24]
    return m_inv @ (np.float32(depth) * np.array([px, py, 1], np.float32) - col)


def fusibile(nd, cams, ref, subset, depth_thresh, normal_thresh, num_consistent, images=None):
    v, h, w, _ = nd.shape
    points, valid = np.zeros((h, w, 12), np.float32), np.zeros((h, w), bool)
    cr = cams[ref]
    for py in range(h):
        for px in range(w):
            normal = nd[ref, py, px]
            X = point_of(cr, px, py, normal[3]).astype(np.float32)
            cx, cn = X.copy(), normal.copy()
            ct = images[ref, py, px].copy() if images is not None else np.zeros(4, np.float32)
            n = 0
            for i in subset:
                if i == ref:
                    continue
                c = cams[i]
                t = c[0:12].reshape(3, 4) @ np.append(X, np.float32(1))
                with np.errstate(all="ignore"):
                    qx, qy = np.float32(t[0] / t[2]), np.float32(t[1] / t[2])
                if not (0 <= qx < w and 0 <= qy < h):
                    continue
                s = tex_linear(nd[i], qx, qy)
                fb = cr[27] * np.sqrt(np.sum((cr[24:27] - c[24:27]) ** 2, dtype=np.float32))
                with np.errstate(all="ignore"):
                    if not abs(fb / np.float32(t[2]) - fb / s[3]) < depth_thresh:
                        continue
                    angle = np.arccos(np.float32(np.dot(s[:3], normal[:3])))
                if np.isnan(angle):
                    angle = 0.0
                if not angle < normal_thresh:
                    continue
                cx += point_of(c, int(qx), int(qy), s[3]).astype(np.float32)
                cn[:3] += s[:3]
                cn[3] = 0
                if images is not None:
                    ct[:3] += tex_linear(images[i], qx, qy)[:3]
                    ct[3] = 0
                n += 1
            if n >= num_consistent:
                valid[py, px] = True
                points[py, px, 0:3], points[py, px, 4:7], points[py, px, 8:11] = cx / (n + 1), cn[:3] / (n + 1), ct[:3] / (n + 1)
    return points, valid
