"""Known-answer tests that pin the restated IEEE-14 data: pypower itself is absent."""
import numpy as np
import torch


def test_kat1_ybus_checksums():
    from rpo_b200.case14 import ybus14
    Y = ybus14()
    G, B = Y.real, Y.imag
    assert abs(G.sum()) < 1e-9
    assert abs(B.sum() - 0.391817) < 5e-6
    assert abs(np.linalg.norm(G) - 27.967487) < 1e-5
    assert abs(np.linalg.norm(B) - 91.015139) < 1e-5
    assert (G != 0).sum() == 42 and (B != 0).sum() == 54
    assert np.abs(Y - Y.T).max() == 0.0
    for (i, j, re, im) in [(0, 0, 6.025029, -19.447071), (0, 1, -4.999132, 15.263086)]:
        assert abs(G[i, j] - re) < 1e-5 and abs(B[i, j] - im) < 1e-5
    assert abs(B[3, 6] - 4.889513) < 1e-5 and abs(B[8, 8] + 24.092506) < 1e-5


def test_kat2_power_flow_reproduces_published_solution(evopf_oracle64):
    """Newton power flow from the case set-points reproduces the solved case printed in the case file itself
    (limited by its 3-decimal precision): validates every bus / branch / generator entry that matters."""
    from rpo_b200.case14 import PD, QD, VA, VM, case14
    orc = evopf_oracle64
    ppc = case14()
    st = torch.zeros(1, 57, dtype=torch.float64)
    st[0, :14] = torch.tensor(ppc["bus"][:, PD] / 100.0)
    st[0, 14:28] = torch.tensor(ppc["bus"][:, QD] / 100.0)
    st[0, 28:33] = 0.2
    z = torch.tensor([[0.4, 0, 0, 0, 1.06, 1.045, 1.01, 1.07, 1.09, 0, 0, 0, 0, 0]], dtype=torch.float64)
    y = orc.complete_partial_fwd(st, z, tol=1e-12)
    vm, va = y[0, 10:24].numpy(), np.rad2deg(y[0, 24:38].numpy())
    assert np.abs(vm - ppc["bus"][:, VM]).max() < 2e-3
    assert np.abs(va - ppc["bus"][:, VA]).max() < 0.03
    assert abs(y[0, 0].item() * 100 - 232.4) < 0.05                 # slack generation, MW
    qg = y[0, 5:10].numpy() * 100
    assert np.abs(qg - np.array([-16.9, 42.4, 23.4, 12.2, 17.4])).max() < 2.0
    assert orc.eq_resid(st, y).abs().max() < 1e-10
