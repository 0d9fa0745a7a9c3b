"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (fp32, NCHW, torch.nn.functional) of the reference's MC-gated PixelCNN as pure functions
over a ``state_dict``-shaped mapping:

    gated_activation   src/models/mcpixelcnn.py:9-20    x, y = chunk(2); mc(relu(bn(x)) * sigmoid(y))
    masked_layer       src/models/mcpixelcnn.py:23-61   vertical / horizontal stacks, causal mask 'A' on layer 0,
                                                        1x1 vert_to_horiz, 1x1 conv -> BN -> MC (+ residual)
    forward            src/models/mcpixelcnn.py:64-101  embedding, layers, 1x1 -> BN -> ReLU -> MC -> 1x1, CE

Pinned against fixtures generated from the unmodified reference (oracle/make_golden_r2.py ->
tests/golden/pixelcnn_small.npz, r2_golden.json["pixelcnn_small"]).
"""
import torch
import torch.nn.functional as F

from .mcgan_oracle import _bn, make_trainable, mc_mask, one_hot  # noqa: F401


def gated_activation(h, ind, sd, pre, training):
    x, y = h.chunk(2, dim=1)
    x = _bn(x, sd, pre + 'bn.', training)
    return mc_mask(F.relu(x) * torch.sigmoid(y), ind, sd[pre + 'mc.codebook'])


def masked_layer(x_v, x_h, ind, sd, pre, training, mask_a, residual):
    wv, wh = sd[pre + 'vert_stack.weight'], sd[pre + 'horiz_stack.weight']
    if mask_a:                                   # make_causal(): in place on the stored weights, every forward
        with torch.no_grad():
            wv[:, :, -1].zero_()
            wh[:, :, :, -1].zero_()
    k = wv.shape[3]
    h_vert = F.conv2d(x_v, wv, sd[pre + 'vert_stack.bias'], padding=(k // 2, k // 2))[:, :, :x_v.size(-1), :]
    out_v = gated_activation(h_vert, ind, sd, pre + 'gate_v.', training)
    h_horiz = F.conv2d(x_h, wh, sd[pre + 'horiz_stack.bias'], padding=(0, k // 2))[:, :, :, :x_h.size(-2)]
    v2h = F.conv2d(h_vert, sd[pre + 'vert_to_horiz.weight'], sd[pre + 'vert_to_horiz.bias'])
    out_h = gated_activation(v2h + h_horiz, ind, sd, pre + 'gate_h.', training)
    r = F.conv2d(out_h, sd[pre + 'horiz_resid.0.module.weight'], sd[pre + 'horiz_resid.0.module.bias'])
    r = mc_mask(_bn(r, sd, pre + 'horiz_resid.1.module.', training), ind, sd[pre + 'horiz_resid.2.codebook'])
    return out_v, (r + x_h if residual else r)


def forward(img, label, sd, num_mode, training=True):
    """img: int64 code map [N, H, W]; -> {'loss', 'logits' [N, num_embedding, H, W]}."""
    ind = one_hot(label, num_mode)
    x = F.embedding(img.reshape(-1), sd['embedding.weight']).view(*img.shape, -1).permute(0, 3, 1, 2).contiguous()
    x_v = x_h = x
    i = 0
    while f'layers.{i}.vert_stack.weight' in sd:
        x_v, x_h = masked_layer(x_v, x_h, ind, sd, f'layers.{i}.', training, mask_a=(i == 0), residual=(i > 0))
        i += 1
    h = F.conv2d(x_h, sd['output_conv.0.module.weight'], sd['output_conv.0.module.bias'])
    h = F.relu(_bn(h, sd, 'output_conv.1.module.', training))
    h = mc_mask(h, ind, sd['output_conv.3.codebook'])
    logits = F.conv2d(h, sd['output_conv.4.module.weight'], sd['output_conv.4.module.bias'])
    return {'loss': F.cross_entropy(logits, img), 'logits': logits}
