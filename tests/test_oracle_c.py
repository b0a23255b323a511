"""-m "not gpu": the C restatement (oracle/rgcn_oracle.c) against the reference's golden vectors."""
import numpy as np

from helpers import layer_case, load_npz, rel_err
from oracle import c_port


def test_c_port_matches_reference_golden():
    blob = load_npz('layer_cases.npz')
    for name in blob['names']:
        c = layer_case(blob, str(name))
        et = c['type_o'] if int(c['reverse']) else c['type_s']
        out = c_port.rgcn_block_layer(c['H'], c['W'], c['Wloop'] if bool(c['self_loop']) else None, c['src'], c['dst'],
                                      et, c['ref_norm'], bool(c['relu']), int(c['nb']))
        assert rel_err(out, c['ref_out']) < 2e-6, name
    out = c_port.rgcn_block_layer(blob['zero_edge/H'], blob['zero_edge/W'], blob['zero_edge/Wloop'],
                                  np.zeros(0), np.zeros(0), np.zeros(0), np.ones(4, np.float32), True, 2)
    assert rel_err(out, blob['zero_edge/ref_out']) < 2e-6
