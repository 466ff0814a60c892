from egopose_b200.nets import FrameContext


class VideoStateNet(FrameContext):
    """models/video_state_net.py:8 constructor signature; on the fused path the context comes from the
    per-frame table (identity v_net), see egopose_b200.nets.FrameContext"""

    def __init__(self, cnn_feat_dim, v_hdim=128, v_margin=10, v_net_type='lstm', v_net_param=None, causal=False):
        super().__init__(cnn_feat_dim, v_hdim, v_margin)
