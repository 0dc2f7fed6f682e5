"""Mirror of the model functions of the reference's RenderNet_Texture_Face_Normal.py -- `decoder_texture` (:34-46),
`RenderNet` (:48-147) -- and of the inference half of its graph wiring (:155-179); BASELINE config 4.
Importable without side effects (the reference trains at import); `is_training` is an explicit argument instead of
the module-global placeholder the reference's RenderNet reads.  Scope names reproduce the reference's, including
its quirks (Image head: block `e_conv7_1` uses scope 'e_conv7_2', default 'conv2d_transpose' scopes; :118-127).
"""
from __future__ import annotations

from . import ops
from . import tfcompat as tf
from .layer_util import (conv2d, conv2d_transpose, conv3d, conv3d_transpose, fully_connected, keep_prob, prelu,
                         projection_unit, res_block_2d, res_block_3d)
from .model_util import tf_transform_voxel_to_match_image
from .resampling_voxel_grid import concat_resampled, tf_rotation_resampling
from .tfcompat import realize


def decoder_texture(z_in):
    """:34-46: texture vector [B,199] -> 3-D texture volume [B,64,64,64,4]."""
    with tf.variable_scope("texture_encoder"):
        batch_size = z_in.shape[0]
        with tf.variable_scope('e_tex_fc1'):
            zP = prelu((fully_connected(z_in, 32 * 32 * 32 * 4)))
            z_resized = realize(zP).reshape(batch_size, 32, 32, 32, 4)
        with tf.variable_scope('e_tex_conv0'):
            conv0 = prelu(conv3d_transpose(z_resized, 4, kernel_size=[4, 4, 4], stride=[1, 1, 1]))
        with tf.variable_scope('e_tex_conv1'):
            conv1 = prelu(conv3d_transpose(conv0, 8, kernel_size=[4, 4, 4], stride=[2, 2, 2]))
        with tf.variable_scope('e_tex_conv2'):
            conv2 = prelu(conv3d(conv1, 4, kernel_size=[4, 4, 4], stride=[1, 1, 1]))
        return conv2


def RenderNet(models_in, prob=0.75, reuse=False, is_training=False):
    """:48-147: 5-channel rotated grid [B,H,W,128,5] -> (albedo image, normal map), each float32 [B,4H,4W,3]."""
    xavier = tf.xavier_initializer
    with tf.variable_scope("encoder"):
        with tf.variable_scope('e_conv1'):
            enc1 = prelu(conv3d(models_in, 8, kernel_size=[5, 5, 5], stride=[2, 2, 2], reuse=reuse, pad="SAME",
                                scope='e_conv1', weight_initializer_type=xavier()))
            enc1 = tf.nn.dropout(enc1, keep_prob(prob, is_training))
        with tf.variable_scope('e_conv2'):
            enc2 = prelu(conv3d(enc1, 16, kernel_size=[3, 3, 3], stride=[1, 1, 2], reuse=reuse, pad="SAME",
                                scope='e_conv2', weight_initializer_type=xavier()))
            enc2 = tf.nn.dropout(enc2, keep_prob(prob, is_training))
        with tf.variable_scope('e_conv3'):
            enc3 = prelu(conv3d(enc2, 16, kernel_size=[3, 3, 3], stride=[1, 1, 1], reuse=reuse, pad="SAME",
                                scope='e_conv3', weight_initializer_type=xavier()))
            enc3 = tf.nn.dropout(enc3, keep_prob(prob, is_training))

        shortcut = enc3
        res = enc3
        for k in range(1, 11):
            res = res_block_3d(res, 16, scope='res1_%d' % k)
        with tf.variable_scope('res1_skip'):
            enc3_skip = conv3d(res, 16, kernel_size=[3, 3, 3], stride=[1, 1, 1], pad="SAME", scope="con1_3X3",
                               weight_initializer_type=xavier())
            enc3_skip = tf.add(tf.cast(enc3_skip, tf.float32), tf.cast(shortcut, tf.float32))

        enc4 = projection_unit(enc3_skip)

        shortcut = enc4
        res = enc4
        for k in range(1, 11):
            res = res_block_2d(res, 32 * 16, scope='res2_%d' % k)
        with tf.variable_scope('res2_skip'):
            enc4_skip = conv2d(res, 32 * 16, kernel_size=[3, 3], stride=[1, 1], scope="con1_3X3",
                               weight_initializer_type=xavier())
            enc4_skip = tf.add(tf.cast(enc4_skip, tf.float32), tf.cast(shortcut, tf.float32))

        with tf.variable_scope('e_conv5'):
            enc5 = prelu(conv2d(enc4_skip, 32 * 8, kernel_size=[4, 4], stride=[1, 1], scope='e_conv5',
                                weight_initializer_type=xavier()))
            enc5 = tf.nn.dropout(enc5, keep_prob(prob, is_training))

        shortcut = enc5
        res = enc5
        for k in range(1, 6):
            res = res_block_2d(res, 32 * 8, scope='res3_%d' % k)
        with tf.variable_scope('res3_skip'):
            enc5_skip = conv2d(res, 32 * 8, kernel_size=[3, 3], stride=[1, 1], scope="con1_3X3",
                               weight_initializer_type=xavier())
            enc5_skip = tf.add(tf.cast(enc5_skip, tf.float32), tf.cast(shortcut, tf.float32))
        enc5_skip = realize(enc5_skip)       # consumed by both heads

        # Two heads with the reference's (irregular) scope names (:113-145):
        #   Image : e_conv6_1/e_conv6_1, e_conv7_1/e_conv7_2 (sic), e_conv8_1|9_1|10_1/conv2d_transpose (default scope)
        #   Normal: e_conv6_2/e_conv6_2, e_conv7_2/e_conv7_2, e_conv8_2/e_conv8_2, e_conv9_2/e_conv9_2, e_conv10_2/e_conv10_2
        heads = (("Image", "1", {"7": "e_conv7_2", "8": None, "9": None, "10": None}),
                 ("Normal", "2", {"7": "e_conv7_2", "8": "e_conv8_2", "9": "e_conv9_2", "10": "e_conv10_2"}))
        outs = []
        for head, sfx, inner in heads:
            def tconv(x, ch, stride, blk):
                kw = dict(weight_initializer_type=xavier())
                if inner[blk] is not None:
                    kw["scope"] = inner[blk]
                return conv2d_transpose(x, ch, [4, 4], stride=[stride, stride], **kw)

            with tf.variable_scope(head):
                with tf.variable_scope('e_conv6_' + sfx):
                    net = prelu(conv2d(enc5_skip, 32 * 4, kernel_size=[4, 4], stride=[1, 1], scope='e_conv6_' + sfx,
                                       weight_initializer_type=xavier()))
                    net = tf.nn.dropout(net, keep_prob(prob, is_training))
                for blk, ch in (("7", 32 * 2), ("8", 32), ("9", 16)):
                    with tf.variable_scope('e_conv%s_%s' % (blk, sfx)):
                        net = prelu(tconv(net, ch, 2, blk))
                        net = tf.nn.dropout(net, keep_prob(prob, is_training))
                with tf.variable_scope('e_conv10_' + sfx):
                    net = tf.nn.sigmoid(tconv(net, 3, 1, "10"), name="encoder_output")
            outs.append(net)
        enc10_1, enc10_2 = outs

        return realize(enc10_1), realize(enc10_2)


def render_graph(model_in, texture_in, param_in, prob=0.75, new_res=128, is_training=False):
    """Inference path of :155-179.  model_in [B,64,64,64,1] ("real_model_in"), texture_in [B,199]
    ("real_texture_in"), param_in [B,3] ("view_name") -> (images_pred, normal_pred)."""
    rotated_models = tf_transform_voxel_to_match_image(tf_rotation_resampling(model_in, param_in, new_size=new_res))
    texture_decoded = decoder_texture(z_in=texture_in)
    texture_rotated = tf_transform_voxel_to_match_image(
        tf_rotation_resampling(texture_decoded, param_in, new_size=new_res))
    model_texture_concat = concat_resampled(rotated_models, texture_rotated)          # tf.concat(..., 4) (:178), kept deferred
    return RenderNet(model_texture_concat, prob=prob, is_training=is_training)
