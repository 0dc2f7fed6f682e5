"""Mirror of the forward model function of the reference's Reconstruct_RenderNet_Face.py that the drop-in boundary lists
(SURVEY §8b): `RenderNet_pretrained(models_in, weight_dict, prob=1.0, trainable=False)` (:113-302) -- the Texture/Normal
RenderNet built from a dictionary of pretrained arrays in the npz-directory key convention of tools/model_util.py:26-39
("e_conv1_e_conv1_weights", "res2_4_con1_3X3_biases", "Image_e_conv6_1_alpha", ...).

It differs from RenderNet_Texture_Face_Normal.RenderNet in exactly the ways the reference's does:
  * the projection unit is written out as reshape + 1x1 `conv2d` under scope `e_conv4` (:168-179);
  * the residual blocks take the `weight_dict` branch of layer_util.res_block_2d/3d, i.e. ReLU instead of PReLU
    (tools/layer_util.py:76,109);
  * head scopes are regular (`Image/e_conv7_1/e_conv7_1`, ...), the last up-convs are `e_conv11_1` and `e_conv11/e_conv11_2`.
Inverse rendering itself (the optimisation loop, :335-537) is out of scope; `rendernet_b200.backward` provides the gradients
of this forward path with respect to its inputs.
"""
from __future__ import annotations

from . import tfcompat as tf
from .layer_util import conv2d, conv2d_transpose, conv3d, prelu, res_block_2d, res_block_3d
from .tfcompat import realize


def RenderNet_pretrained(models_in, weight_dict, prob=1.0, trainable=False):
    """models_in: rotated + axis-transformed 5-channel grid [B,H,W,128,5] (geometry + 4 texture channels; the reference's
    docstring says 64^3 x 6, its graph feeds the 128^3 x 5 concat of :360-378).  Returns (albedo, normal), float32 [B,4H,4W,3]."""
    if float(prob) != 1.0:
        raise NotImplementedError("inference path: dropout keep probability must be 1.0")
    wd = weight_dict

    def conv_args(key):
        return dict(trainable=trainable, weight_initializer=wd[key + "_weights"], bias_initializer=wd[key + "_biases"])

    with tf.variable_scope("encoder"):
        net = models_in
        for name, ch, k, stride in (("e_conv1", 8, 5, [2, 2, 2]), ("e_conv2", 16, 3, [1, 1, 2]), ("e_conv3", 16, 3, [1, 1, 1])):
            with tf.variable_scope(name):
                net = prelu(conv3d(net, ch, kernel_size=[k, k, k], stride=stride, pad="SAME", scope=name,
                                   **conv_args(f"{name}_{name}")), alpha=wd[name + "_alpha"], trainable=trainable)
        shortcut = net
        for i in range(1, 11):
            net = res_block_3d(net, 16, scope='res1_%d' % i, weight_dict=wd, trainable=trainable)
        with tf.variable_scope('res1_skip'):
            skip = conv3d(net, 16, kernel_size=[3, 3, 3], stride=[1, 1, 1], pad="SAME", scope="con1_3X3",
                          **conv_args("res1_skip_con1_3X3"))
            enc3_skip = realize(tf.add(tf.cast(skip, tf.float32), tf.cast(shortcut, tf.float32)))

        # collapse the depth axis (free in channel-last) and mix it with a 1x1 convolution (:168-179)
        B, H, W, D, C = enc3_skip.shape
        enc3_2d = enc3_skip.reshape(B, H, W, D * C)
        with tf.variable_scope('e_conv4'):
            net = prelu(conv2d(enc3_2d, num_outputs=D * C, kernel_size=[1, 1], scope='e_conv4', **conv_args("e_conv4_e_conv4")),
                        alpha=wd["e_conv4_alpha"], trainable=trainable)

        for stage, width, nblocks, nxt, nxt_ch in (("res2", 32 * 16, 10, "e_conv5", 32 * 8), ("res3", 32 * 8, 5, None, 0)):
            shortcut = net
            for i in range(1, nblocks + 1):
                net = res_block_2d(net, width, scope='%s_%d' % (stage, i), weight_dict=wd, trainable=trainable)
            with tf.variable_scope(stage + '_skip'):
                skip = conv2d(net, width, kernel_size=[3, 3], scope="con1_3X3", **conv_args(stage + "_skip_con1_3X3"))
                net = tf.add(tf.cast(skip, tf.float32), tf.cast(shortcut, tf.float32))
            if nxt is not None:
                with tf.variable_scope(nxt):
                    net = prelu(conv2d(net, nxt_ch, kernel_size=[4, 4], scope=nxt, **conv_args(f"{nxt}_{nxt}")),
                                alpha=wd[nxt + "_alpha"], trainable=trainable)
        enc5_skip = realize(net)                           # consumed by both heads

        outs = []
        for head, sfx, last_outer in (("Image", "1", "e_conv11_1"), ("Normal", "2", "e_conv11")):
            with tf.variable_scope(head):
                name = "e_conv6_" + sfx
                with tf.variable_scope(name):
                    net = prelu(conv2d(enc5_skip, 32 * 4, kernel_size=[4, 4], scope=name, **conv_args(f"{head}_{name}_{name}")),
                                alpha=wd[f"{head}_{name}_alpha"], trainable=trainable)
                for blk, ch in (("7", 32 * 2), ("8", 32), ("9", 16)):
                    name = "e_conv%s_%s" % (blk, sfx)
                    with tf.variable_scope(name):
                        net = prelu(conv2d_transpose(net, ch, [4, 4], stride=[2, 2], scope=name,
                                                     **conv_args(f"{head}_{name}_{name}")),
                                    alpha=wd[f"{head}_{name}_alpha"], trainable=trainable)
                name = "e_conv11_" + sfx
                with tf.variable_scope(last_outer):
                    net = tf.nn.sigmoid(conv2d_transpose(net, 3, [4, 4], stride=[1, 1], scope=name,
                                                         **conv_args(f"{head}_{name}_{name}")), name="encoder_output")
            outs.append(realize(net))
    return outs[0], outs[1]


def pretrained_dict_from_texture_weights(W):
    """Re-key a Texture/Normal weight dict in TF variable naming (RenderNet_Texture_Face_Normal scopes, e.g. the oracle's
    `init_texture_weights`) into the npz-directory keys `RenderNet_pretrained` reads.  Both functions then compute the same
    network provided the residual-block alphas are zero (the weight_dict branch uses ReLU).  Test / migration helper."""
    out = {}
    ren = {"projection_unit/Conv": "e_conv4/e_conv4", "projection_unit/alpha": "e_conv4/alpha",
           "Image/e_conv7_1/e_conv7_2": "Image/e_conv7_1/e_conv7_1", "Image/e_conv8_1/conv2d_transpose": "Image/e_conv8_1/e_conv8_1",
           "Image/e_conv9_1/conv2d_transpose": "Image/e_conv9_1/e_conv9_1",
           "Image/e_conv10_1/conv2d_transpose": "Image/e_conv11_1/e_conv11_1", "Normal/e_conv10_2/e_conv10_2": "Normal/e_conv11_2/e_conv11_2"}
    for k, v in W.items():
        if not k.startswith("encoder/"):
            continue
        name = k[len("encoder/"):]
        for a, b in ren.items():
            if name.startswith(a):
                name = b + name[len(a):]
                break
        out[name.replace("/", "_")] = v
    return out
