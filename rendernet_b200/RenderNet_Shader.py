"""Mirror of the reference's RenderNet_Shader.py model function (`RenderNet`, :32-131) and of the inference
half of its graph wiring (:139-156), importable without side effects (the reference module trains at import
and reads the global `cfg`; here `is_greyscale` is an explicit argument).

The body below is written against `layer_util` / `tfcompat` exactly the way the reference is written against
tools/layer_util.py / TensorFlow, so scopes -- and therefore variable names -- are identical:
encoder/e_conv1/e_conv1/weights, encoder/res2_3/con1_3X3/biases, encoder/projection_unit/Conv/weights, ...
"""
from __future__ import annotations

from . import tfcompat as tf
from .layer_util import conv3d, keep_prob, prelu, projection_unit, res_block_2d, res_block_3d, slim
from .model_util import tf_transform_voxel_to_match_image
from .resampling_voxel_grid import tf_rotation_resampling
from .tfcompat import realize


def RenderNet(models_in, is_training=False, prob=0.75, reuse=False, is_greyscale=False, stages=None):
    """3-D encoder -> projection unit -> 2-D residual trunk -> up-conv decoder -> sigmoid image.
    models_in [B,H,W,128,1] (resampled + axis-transformed grid); returns float32 [B,4H,4W,3|1].
    `stages` (optional dict) receives the realised stage tensors named as in the oracle, for parity tests."""
    xavier = tf.xavier_initializer

    def keep(name, t):
        if stages is not None:
            stages[name] = realize(t)
        return t

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
            enc3 = prelu(conv3d(enc2, 32, kernel_size=[3, 3, 3], stride=[1, 1, 1], reuse=reuse, pad="SAME",
                                scope='e_conv3', weight_initializer_type=xavier()))
            enc3 = tf.nn.dropout(enc3, keep_prob(prob, is_training))

        shortcut = keep('enc3', enc3)
        res = enc3
        for k in range(1, 11):                                           # res1_1 .. res1_10 (:51-60)
            res = res_block_3d(res, 32, scope='res1_%d' % k)

        with tf.variable_scope('res1_skip'):
            enc3_skip = conv3d(res, 32, kernel_size=[3, 3, 3], stride=[1, 1, 1], pad="SAME", scope="con1_3X3",
                               weight_initializer_type=xavier())
            enc3_skip = tf.add(tf.cast(enc3_skip, tf.float32), tf.cast(shortcut, tf.float32))

        keep('enc3_skip', enc3_skip)
        enc4 = projection_unit(enc3_skip)
        keep('enc4', enc4)

        shortcut = enc4
        res = enc4
        for k in range(1, 11):                                           # res2_1 .. res2_10 (:71-80)
            res = res_block_2d(res, 32 * 32, scope='res2_%d' % k)

        with tf.variable_scope('res2_skip'):
            enc4_skip = slim.conv2d(res, 32 * 32, kernel_size=3, stride=1, activation_fn=None, scope="con1_3X3")
            enc4_skip = tf.add(tf.cast(enc4_skip, tf.float32), tf.cast(shortcut, tf.float32))

        keep('enc4_skip', enc4_skip)
        with tf.variable_scope('e_conv5'):
            enc5 = prelu(slim.conv2d(inputs=enc4_skip, num_outputs=32 * 16, kernel_size=(4, 4), stride=1,
                                     activation_fn=None, scope='e_conv5'))
            enc5 = tf.nn.dropout(enc5, keep_prob(prob, is_training))
        shortcut = enc5
        res = enc5
        for k in range(1, 6):                                            # res3_1 .. res3_5 (:91-95)
            res = res_block_2d(res, 32 * 16, scope='res3_%d' % k)

        with tf.variable_scope('res3_skip'):
            enc5_skip = slim.conv2d(res, 32 * 16, kernel_size=3, stride=1, activation_fn=None, scope="con1_3X3")
            enc5_skip = tf.add(tf.cast(enc5_skip, tf.float32), tf.cast(shortcut, tf.float32))

        keep('enc5_skip', enc5_skip)
        with tf.variable_scope('e_conv6'):
            enc6 = prelu(slim.conv2d(inputs=enc5_skip, num_outputs=32 * 8, kernel_size=(4, 4), stride=1,
                                     activation_fn=None, scope='e_conv6'))
            enc6 = tf.nn.dropout(enc6, keep_prob(prob, is_training))

        # e_conv7 .. e_conv10: 4x4 transposed convs (stride 2,1,2,2,1) + PReLU; variable scopes <name>/<name> (:105-123)
        net = enc6
        for name, ch, stride in (('e_conv7', 32 * 4, 2), ('e_conv7_1', 32 * 4, 1), ('e_conv8', 32 * 2, 2),
                                 ('e_conv9', 32, 2), ('e_conv10', 16, 1)):
            with tf.variable_scope(name):
                net = prelu(slim.conv2d_transpose(net, ch, (4, 4), stride=stride, activation_fn=None, scope=name))
                net = tf.nn.dropout(net, keep_prob(prob, is_training))
        enc10 = net
        keep('enc10', enc10)
        n_out = 1 if is_greyscale else 3                                 # cfg['is_greyscale'] (:125)
        enc11 = slim.conv2d_transpose(enc10, n_out, (4, 4), stride=1, activation_fn=None, scope='e_conv11')
        output = tf.nn.sigmoid(enc11, name="output")
        return realize(output)


def render_graph(model_in, param_in, is_training=False, prob=0.75, new_res=128, is_greyscale=False):
    """Inference path of the reference graph (:139-156): rotate/resample -> axis transform -> RenderNet.
    model_in [B,64,64,64,1] fp32 ("real_model_in"), param_in [B,3] ("view_name")."""
    rotated_models = tf_rotation_resampling(model_in, param_in, new_size=new_res)
    rotated_models = tf_transform_voxel_to_match_image(rotated_models)
    return RenderNet(models_in=rotated_models, is_training=is_training, prob=prob, is_greyscale=is_greyscale)
