"""Test helpers: convert the engine's internal layouts (debug taps) to the reference's NCHW / OIHW."""
import numpy as np
import torch


def geometry(args):
    h, w, cin = int(args.image_height), int(args.image_width), int(args.image_channels)
    geo = []
    for _ in range(int(args.num_stages)):
        geo.append(dict(h=h, w=w, cin=cin))
        h, w, cin = h // 2, w // 2, int(args.cnn_num_filters)
    return geo, (h, w)


def grid_to_nchw(buf, n, h, w, F):
    """[n*(h+1)*(w+1), F] padded pixel grid (row 0 / column 0 of every image block are the shared zero padding)
    -> [n, F, h, w]."""
    a = np.asarray(buf).reshape(n, h + 1, w + 1, F)[:, 1:h + 1, 1:w + 1, :]
    return torch.from_numpy(np.ascontiguousarray(a.transpose(0, 3, 1, 2)))


def flat_to_nchw(buf, n, h, w, F):
    """[n, h*w, F] unpadded (last block's pooled output / features) -> [n, F, h, w]."""
    a = np.asarray(buf).reshape(n, h, w, F)
    return torch.from_numpy(np.ascontiguousarray(a.transpose(0, 3, 1, 2)))


def theta_to_ref(vec, args):
    """Internal fast-weight vector -> {reference name: tensor in reference layout}."""
    geo, (ph, pw) = geometry(args)
    F, N = int(args.cnn_num_filters), int(args.num_classes_per_set)
    out, o = {}, 0
    v = np.asarray(vec)
    for l, g in enumerate(geo):
        cin = g["cin"]
        wsz = 9 * cin * F
        w = v[o:o + wsz].reshape(3, 3, cin, F).transpose(3, 2, 0, 1)     # [tap(ky,kx)][c][f] -> [f][c][ky][kx]
        out["classifier.layer_dict.conv%d.conv.weight" % l] = torch.from_numpy(np.ascontiguousarray(w))
        o += wsz
        out["classifier.layer_dict.conv%d.conv.bias" % l] = torch.from_numpy(v[o:o + F].copy())
        o += F
    pix = ph * pw
    D = pix * F
    fw = v[o:o + N * D].reshape(N, pix, F).transpose(0, 2, 1).reshape(N, D)   # [k][pix][c] -> [k][c*pix + pix]
    out["classifier.layer_dict.linear.weights"] = torch.from_numpy(np.ascontiguousarray(fw))
    o += N * D
    out["classifier.layer_dict.linear.bias"] = torch.from_numpy(v[o:o + N].copy())
    return out


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
