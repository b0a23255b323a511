"""Backward of the fused read-out + GRU (placeholder until renet_gru_bwd lands)."""


def fused_gru_backward(ctx, dhn4, dhn3):
    raise NotImplementedError('renet_b200: backward through the fused GRU is not implemented yet; '
                              'use RENet.forward_unfused for training')
