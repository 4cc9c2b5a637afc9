"""GPU parity of the policy-network kernels (through the C ABI / torch ops) against logits produced by
the unmodified reference on the shipped weights (tests/golden) and against the torch-CPU oracle.
fp32: logits within 1e-4 relative (north star); chosen actions bit-exact wherever the reference's own
top-2 margin exceeds the fp32 noise floor."""
import numpy as np
import pytest
import torch

from helpers import load_network, model_config, oracle_params, rel_err, top2_margin
from oracle import env_oracle as eo
from oracle import model_oracle as mo

pytestmark = pytest.mark.gpu
REL = 1e-4
MARGIN_FLOOR = 1e-4   # decisions closer than this in the reference may legitimately flip


def _check_actions(mine, ref_logits):
    ref_act = ref_logits.argmax(-1)
    bad = mine != ref_act
    if bad.any():
        assert top2_margin(ref_logits)[bad].max() < MARGIN_FLOOR, "action flipped where the reference margin is wide"
    return int(bad.sum())


@pytest.mark.parametrize("name", ["gnn_k3", "gnn_k2", "gnn_msg_k3", "baseline", "gnn_k4_seed4"])
def test_logits_match_reference_goldens(model_logits, name):
    states = torch.from_numpy(model_logits[f"{name}/states"]).float().cuda()
    gso = torch.from_numpy(model_logits[f"{name}/gso"]).float().cuda()
    net = load_network(name, num_agents=states.shape[1]).eval()
    with torch.no_grad():
        logits = net(states) if name == "baseline" else net(states, gso)
        logits2, actions = net.act(states, None if name == "baseline" else gso)
    ref = model_logits[f"{name}/logits"]
    assert tuple(logits.shape) == ref.shape
    assert rel_err(logits.cpu().numpy(), ref) < REL
    assert torch.equal(logits, logits2)
    assert _check_actions(actions.cpu().numpy(), ref) == 0
    assert np.array_equal(actions.cpu().numpy(), logits.cpu().numpy().argmax(-1))


@pytest.mark.parametrize("name", ["gcn_k3_n5", "msg_k3_n5", "gcn_k4_n20", "msg_k2_n7"])
def test_graph_layers_match_reference_goldens(graph_layers, name):
    import mapf_gnn_b200 as m
    g = graph_layers
    x = torch.from_numpy(g[f"{name}/x"]).cuda()
    adj = torch.from_numpy(g[f"{name}/adj"]).float().cuda()
    W = g[f"{name}/W"]
    F, K, G = W.shape
    cls = m.MessagePassingLayer if f"{name}/W2" in g.files else m.GCNLayer
    layer = cls(x.shape[2], F, G, K).cuda()
    with torch.no_grad():
        layer.W.copy_(torch.from_numpy(W))
        if f"{name}/W2" in g.files:
            layer.W2.copy_(torch.from_numpy(g[f"{name}/W2"]))
        layer.addGSO(adj)
        y = layer(x)
    assert rel_err(y.cpu().numpy(), g[f"{name}/y"]) < 1e-5


@pytest.mark.parametrize("name,B,N", [("gnn_k3", 300, 5), ("gnn_msg_k3", 67, 20), ("gnn_k4_seed4", 9, 64),
                                      ("baseline", 257, 5), ("gnn_k2", 1, 5), ("gnn_k3", 33, 13)])
def test_logits_match_oracle_on_random_observations(name, B, N):
    rng = np.random.default_rng(B * 31 + N)
    H = 12 if N <= 20 else 24
    starts, goals, obst = eo.make_scenarios(rng, B, N, H, 6)
    S, A = [], []
    for b in range(B):
        env = eo.EnvOracle(N, H, goals[b], starts[b], obst[b], sensing_range=4)
        o, _ = env.step(rng.integers(0, 5, size=N))
        S.append(o["fov"]), A.append(o["adj_matrix"])
    states = torch.tensor(np.stack(S)).float()
    gso = torch.tensor(np.stack(A)).float()
    with torch.no_grad():
        ref = mo.forward(oracle_params(name), states, gso).numpy()
    net = load_network(name, num_agents=N).eval()
    logits, actions = net.act(states.cuda(), None if name == "baseline" else gso.cuda())
    assert rel_err(logits.cpu().numpy(), ref) < REL
    _check_actions(actions.cpu().numpy(), ref)


@pytest.mark.parametrize("rows", [1, 15, 16, 17, 33, 640, 2368, 2400, 4736 + 300, 4736 + 74 * 32, 4736 + 75 * 32 + 5, 2 * 4736 + 16])
def test_encoder_tile_schedules(rows):
    """The persistent encoder walks full 32-image tiles and, for a ragged last round, 16-image half tiles (different
    lane mapping); every schedule must give the oracle's features (framework_gnn.py:156-166), row for row."""
    rng = np.random.default_rng(rows)
    states = torch.tensor(rng.integers(0, 4, size=(rows, 1, 2, 5, 5))).float()
    states[:: 7] *= 0.37                                   # not only small integers
    params = oracle_params("gnn_k3")
    with torch.no_grad():
        ref = mo.encode_agents(params, states)[:, :, 0].numpy()
    net = load_network("gnn_k3", num_agents=1).eval()
    x = torch.ops.mapf.encoder_fwd(net.dims(), states.cuda(), net.packed_weights()).cpu().numpy()
    assert x.shape == ref.shape
    assert rel_err(x, ref) < REL
    assert np.abs(x - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("rows", [1, 17, 129, 640, 2400, 4736 + 300, 4736 + 75 * 32 + 5])
def test_presplit_tensor_core_fc_matches_oracle(rows):
    """Conv kernel emitting the pre-split / pre-tiled A operand + all-TMA tcgen05 FC (mapf_encoder_conv_split_fwd,
    mapf_tc_fc_presplit) against the oracle encoder, over full-tile, half-tile and ragged schedules."""
    from mapf_gnn_b200 import _lib
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    rng = np.random.default_rng(rows + 1)
    states = torch.tensor(rng.integers(0, 4, size=(rows, 1, 2, 5, 5))).float()
    states[:: 5] *= 0.61
    params = oracle_params("gnn_k3")
    with torch.no_grad():
        ref = mo.encode_agents(params, states)[:, :, 0].numpy()
    net = load_network("gnn_k3", num_agents=1).eval()
    dims, packed = net.dims(), net.packed_weights()
    p = net_params(dims)
    st = states.cuda()
    Ap = torch.full((lib.mapf_tc_presplit_a_size(rows, dims[6]),), float("nan"), device="cuda")
    x = torch.empty((rows, dims[7]), device="cuda")
    fcw = net.compressMLP[0].weight.detach().contiguous()
    fcb = net.compressMLP[0].bias.detach().contiguous()
    Bp = torch.empty((lib.mapf_tc_packed_b_size(dims[7], dims[6]),), device="cuda")
    s = _lib.current_stream()
    _lib.check(lib.mapf_tc_pack_b(fcw.data_ptr(), dims[6], None, 0, dims[7], Bp.data_ptr(), s), "pack")
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    _lib.check(lib.mapf_encoder_conv_split_fwd(p, ea, s), "conv_split")
    _lib.check(lib.mapf_tc_fc_presplit(rows, dims[7], dims[6], Ap.data_ptr(), Bp.data_ptr(), x.data_ptr(), dims[7],
                                       fcb.data_ptr(), 1, s), "fc_presplit")
    got = x.cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_err(got, ref) < REL


def _untile_presplit_f16(Ap, rows, K):
    """enc_tc.cu output: fp16 hi|lo tiles [row/128][k/16][hi|lo][(row%128)/8][(k%16)/8][row%8][k%8] followed by the
    rows' inverse scales -> (hi + lo) * inverse scale as [rows, K] float64."""
    nblk, nch = (rows + 127) // 128, (K + 15) // 16
    words = nblk * nch * 2048
    h = Ap[:words].view(torch.float16).cpu().numpy().astype(np.float64)
    a = h.reshape(nblk, nch, 2, 16, 2, 8, 8).transpose(2, 0, 3, 5, 1, 4, 6).reshape(2, nblk * 128, nch * 16)
    inv = Ap[words:words + rows].cpu().numpy().astype(np.float64)
    return (a[0] + a[1])[:rows, :K] * inv[:, None]


def _conv_stack_reference(params, states, depth):
    import torch.nn.functional as F
    h = states
    for l in range(depth):
        li = 3 * l
        h = F.conv2d(h, params[f"convLayers.{li}.weight"], params[f"convLayers.{li}.bias"], padding=1)
        bn = f"convLayers.{li + 1}"
        h = F.relu(F.batch_norm(h, params[f"{bn}.running_mean"], params[f"{bn}.running_var"], params[f"{bn}.weight"],
                                params[f"{bn}.bias"], False, 0.1, 1e-5))
    return h.permute(0, 2, 3, 1).reshape(states.shape[0], -1).numpy()     # pixel-major (y, x, c)


@pytest.mark.parametrize("depth", [1, 2, 3])
@pytest.mark.parametrize("rows", [1, 3, 4, 5, 127, 129, 1777, 4736 + 300])
def test_tensor_core_conv_stack_matches_torch(rows, depth):
    """enc_tc.cu: conv + BN + ReLU layers as tcgen05 GEMMs on fp16-split operands with per-image dynamic scales
    (framework_gnn.py:55-71,160-163).  The emitted FC operand, re-assembled ((hi + lo) / scale), must equal the torch-CPU conv stack
    to fp32 accuracy for every tile raggedness and depth -- including all-zero images, huge, tiny, negative and
    non-integer observations (the scale logic), which the FOV tensors of a rollout never contain."""
    from mapf_gnn_b200 import _lib
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    rng = np.random.default_rng(1000 * depth + rows)
    states = torch.tensor(rng.integers(0, 4, size=(rows, 2, 5, 5))).float()
    states[1::5] *= 0.61
    states[2::7] = torch.randn_like(states[2::7])
    states[3::11] = 0.0
    states[4::13] *= 3.0e4
    states[5::17] *= 1.0e-6
    params = oracle_params("gnn_k3")
    with torch.no_grad():
        ref = _conv_stack_reference(params, states, depth)
    net = load_network("gnn_k3", num_agents=1).eval()
    g = net.param_groups()
    d = lambda ts: [t.detach() for t in ts[:depth]]
    dims = [depth, 2] + [16] * depth + [0] * (4 - depth) + [400, 64, 0, 0, 5]
    packed = torch.ops.mapf.pack_weights(dims, d(g["conv_w"]), d(g["conv_b"]), d(g["bn_gamma"]), d(g["bn_beta"]),
                                         d(g["bn_mean"]), d(g["bn_var"]), g["fc_w"].detach(), g["fc_b"].detach(), None,
                                         None, torch.zeros(5, 64, device="cuda"), torch.zeros(5, device="cuda"))
    p = net_params(dims)
    assert lib.mapf_encoder_tc_supported(p) == 1
    Ap = torch.full((lib.mapf_tc_presplit_a_size(rows, 400),), float("nan"), device="cuda")
    st = states.cuda()
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    _lib.check(lib.mapf_encoder_tc_conv_fwd(p, ea, _lib.current_stream()), "tc_conv")
    got = _untile_presplit_f16(Ap, rows, 400)
    assert np.isfinite(got).all()
    # bounds canary: behind the operand tiles and the row scales the NaN fill of the workspace must be untouched
    used = ((rows + 127) // 128) * 25 * 2048 + rows
    assert bool(torch.isnan(Ap[used:]).all())
    # per image: 1e-5 of the image's largest activation (the split keeps 22 bits relative to it)
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-30)
    assert (np.abs(got - ref) / scale).max() < 1e-5


@pytest.mark.parametrize("rows", [1, 5, 128, 131, 2400, 4736 + 300])
def test_tensor_core_encoder_matches_oracle(rows):
    """Conv stack on the tensor cores + fp16-split FC (mapf_encoder_tc_conv_fwd, mapf_tc_pack_b_f16_pixel_major,
    mapf_tc_fc_presplit_f16) against the oracle encoder (framework_gnn.py:156-166), through the C ABI."""
    from mapf_gnn_b200 import _lib
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    rng = np.random.default_rng(rows + 7)
    states = torch.tensor(rng.integers(0, 4, size=(rows, 1, 2, 5, 5))).float()
    states[:: 5] *= 0.61
    params = oracle_params("gnn_k3")
    with torch.no_grad():
        ref = mo.encode_agents(params, states)[:, :, 0].numpy()
    net = load_network("gnn_k3", num_agents=1).eval()
    dims, packed = net.dims(), net.packed_weights()
    p = net_params(dims)
    st = states.cuda()
    Ap = torch.full((lib.mapf_tc_presplit_a_size(rows, dims[6]),), float("nan"), device="cuda")
    xg = torch.full((rows + 8, dims[7]), float("nan"), device="cuda")     # 8 canary rows behind the output
    x = xg[:rows]
    fcw = net.compressMLP[0].weight.detach().contiguous()
    fcb = net.compressMLP[0].bias.detach().contiguous()
    Bp = torch.empty((lib.mapf_tc_packed_b_f16_size(dims[7], dims[6]),), device="cuda")
    s = _lib.current_stream()
    _lib.check(lib.mapf_tc_pack_b_f16_pixel_major(fcw.data_ptr(), 16, dims[7], Bp.data_ptr(), s), "pack")
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    _lib.check(lib.mapf_encoder_tc_conv_fwd(p, ea, s), "tc_conv")
    _lib.check(lib.mapf_tc_fc_presplit_f16(rows, dims[7], dims[6], Ap.data_ptr(), Bp.data_ptr(), x.data_ptr(), dims[7],
                                           fcb.data_ptr(), 1, s), "fc_f16")
    got = x.cpu().numpy()
    assert np.isfinite(got).all()
    assert bool(torch.isnan(xg[rows:]).all())
    assert rel_err(got, ref) < 1e-5


@pytest.mark.parametrize("depth", [2, 3])
@pytest.mark.parametrize("rows", [1, 7, 12, 13, 131, 1777 + 5])
def test_tensor_core_conv_stack_baseline_widths(rows, depth):
    """The baseline network's conv stack (first layer 2 -> 32 channels, framework_baseline.py): layer 0 runs as two
    16-channel GEMM rounds, layer 1 contracts over 3 x 32 operand entries, three warpgroups per CTA."""
    from mapf_gnn_b200 import _lib
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    rng = np.random.default_rng(77 * depth + rows)
    states = torch.tensor(rng.integers(0, 4, size=(rows, 2, 5, 5))).float()
    states[1::5] *= 0.61
    states[2::7] = torch.randn_like(states[2::7])
    states[3::11] = 0.0
    params = oracle_params("baseline")
    with torch.no_grad():
        ref = _conv_stack_reference(params, states, depth)
    net = load_network("baseline", num_agents=1).eval()
    g = net.param_groups()
    d = lambda ts: [t.detach() for t in ts[:depth]]
    dims = [depth, 2, 32] + [16] * (depth - 1) + [0] * (4 - depth) + [400, 64, 0, 0, 5]
    packed = torch.ops.mapf.pack_weights(dims, d(g["conv_w"]), d(g["conv_b"]), d(g["bn_gamma"]), d(g["bn_beta"]),
                                         d(g["bn_mean"]), d(g["bn_var"]), g["fc_w"].detach(), g["fc_b"].detach(), None,
                                         None, torch.zeros(5, 64, device="cuda"), torch.zeros(5, device="cuda"))
    p = net_params(dims)
    assert lib.mapf_encoder_tc_supported(p) == 1
    Ap = torch.full((lib.mapf_tc_presplit_a_size(rows, 400),), float("nan"), device="cuda")
    st = states.cuda()
    ea = _lib.EncoderFwdArgs(rows, st.data_ptr(), packed.data_ptr(), Ap.data_ptr())
    _lib.check(lib.mapf_encoder_tc_conv_fwd(p, ea, _lib.current_stream()), "tc_conv")
    got = _untile_presplit_f16(Ap, rows, 400)
    assert np.isfinite(got).all()
    used = ((rows + 127) // 128) * 25 * 2048 + rows
    assert bool(torch.isnan(Ap[used:]).all())
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-30)
    assert (np.abs(got - ref) / scale).max() < 1e-5


def test_tensor_core_encoder_width_support():
    """The tensor-core conv stack takes a first layer of 16 or 32 channels with 16-channel layers behind it; anything
    else must be reported unsupported (the policy then keeps the FFMA2 encoder) instead of being mis-computed."""
    from mapf_gnn_b200 import _lib
    from mapf_gnn_b200.ops import net_params
    lib = _lib.load()
    assert lib.mapf_encoder_tc_supported(net_params(load_network("baseline").dims())) == 1
    assert lib.mapf_encoder_tc_supported(net_params(load_network("gnn_k2").dims())) == 1
    assert lib.mapf_encoder_tc_supported(net_params([3, 2, 16, 32, 16, 0, 400, 64, 0, 0, 5])) == 0
    assert lib.mapf_encoder_tc_supported(net_params([1, 2, 32, 0, 0, 0, 800, 64, 0, 0, 5])) == 0


def test_programmatic_dependent_launch_does_not_change_results(monkeypatch):
    """The step kernels are chained by programmatic dependent launch (griddepcontrol); with the launch attribute switched
    off (mapf_policy_fwd_args.flags & MAPF_POLICY_NO_PDL; the Python layer maps MAPF_B200_NO_PDL=1 to it per call) a
    rollout must give bit-identical logits, actions and positions: the early prologues may only touch data that is
    older than the kernel in front."""
    import mapf_gnn_b200 as m

    def run(name, N):
        rng = np.random.default_rng(5)
        starts, goals, obst = m.make_scenarios(rng, 64, N, 14, 6)
        env = m.BatchedGraphEnv(model_config(name, N, board_size=[14, 14]), goals, starts, obst, sensing_range=5)
        ro = m.Rollout(load_network(name, num_agents=N).eval(), env)
        ro.run(7)
        torch.cuda.synchronize()
        return {"logits": ro.logits.cpu().numpy(), "pos": env.pos.cpu().numpy(), "actions": ro.actions.cpu().numpy(),
                "fov": env.fov.cpu().numpy(), "adj": env.adj_matrix.cpu().numpy()}

    for name, N in (("gnn_k3", 5), ("baseline", 5), ("gnn_msg_k3", 20)):
        monkeypatch.delenv("MAPF_B200_NO_PDL", raising=False)
        a = run(name, N)
        monkeypatch.setenv("MAPF_B200_NO_PDL", "1")
        b = run(name, N)
        for k in a:
            assert np.array_equal(a[k], b[k]), (name, k)


def test_training_gso_convention_a_plus_i():
    """The data loader feeds A + I (data_loader.py:74); GCNLayer adds I again -> diagonal 2."""
    rng = np.random.default_rng(3)
    B, N = 16, 5
    states = torch.tensor(rng.integers(0, 4, size=(B, N, 2, 5, 5))).float()
    adj = torch.tensor((rng.random((B, N, N)) < 0.4).astype(np.float32))
    gso = adj + torch.eye(N)
    for name in ("gnn_k3", "gnn_msg_k3"):
        with torch.no_grad():
            ref = mo.forward(oracle_params(name), states, gso).numpy()
        net = load_network(name).eval()
        with torch.no_grad():
            out = net(states.cuda(), gso.cuda())
        assert rel_err(out.cpu().numpy(), ref) < REL


def test_weights_repack_after_update():
    net = load_network("gnn_k3").eval()
    states = torch.rand(4, 5, 2, 5, 5, device="cuda")
    gso = torch.zeros(4, 5, 5, device="cuda")
    with torch.no_grad():
        a = net(states, gso).clone()
        net.actionMLP[0].bias.add_(1.0)
        b = net(states, gso)
    assert torch.allclose(b, a + 1.0, atol=1e-5)


# name, E, N, H, M, T, seed: gnn_k3 at the C1 shape, the baseline at the C2b shape, the message-passing network at the C3
# shape (N = 20: separate env kernel behind the fused graph kernel), K = 4 at N = 64 (taps + mix kernels, C5 shape)
ROLLOUT_CASES = [("gnn_k3", 48, 5, 18, 5, 12, 0), ("baseline", 48, 5, 28, 8, 12, 1), ("gnn_k2", 48, 5, 28, 8, 12, 1),
                 ("gnn_msg_k3", 24, 20, 32, 40, 10, 2), ("gnn_k4_seed4", 6, 64, 64, 200, 5, 4)]


@pytest.mark.parametrize("no_pdl", [False, True])
@pytest.mark.parametrize("name,E,N,H,M,T,seed", ROLLOUT_CASES)
def test_rollout_matches_teacher_forced_oracle(name, E, N, H, M, T, seed, no_pdl, monkeypatch):
    """Full on-device rollout (policy -> argmax -> env step) vs oracle env + oracle model, teacher-forced:
    the oracle follows the device's actions; device actions must equal the oracle's argmax wherever the
    oracle's margin is above the fp32 noise floor, and states must stay bit-identical.  Every network family at its
    BASELINE shape, with and without programmatic dependent launch (mapf_policy_fwd_args.flags & MAPF_POLICY_NO_PDL)."""
    import mapf_gnn_b200 as m
    if no_pdl:
        monkeypatch.setenv("MAPF_B200_NO_PDL", "1")
    rng = np.random.default_rng(seed)
    starts, goals, obst = eo.make_scenarios(rng, E, N, H, M)
    cfg = model_config(name, N, board_size=[H, H])
    env = m.BatchedGraphEnv(cfg, goals, starts, obst, sensing_range=6)
    net = load_network(name, num_agents=N).eval()
    ro = m.Rollout(net, env)
    params = oracle_params(name)
    oracles = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e], sensing_range=6) for e in range(E)]
    obs = [o.observe() for o in oracles]
    alive = np.ones(E, dtype=bool)
    flips = 0
    for t in range(T):
        ro.run(1)
        acts = ro.actions.cpu().numpy()
        states = torch.tensor(np.stack([o["fov"] for o in obs])).float()
        gso = torch.tensor(np.stack([o["adj_matrix"] for o in obs])).float()
        with torch.no_grad():
            ref = mo.forward(params, states, gso).numpy()
        assert rel_err(ro.logits.cpu().numpy()[alive], ref[alive]) < REL
        flips += _check_actions(acts[alive], ref[alive])
        pos, fov, adj = env.pos.cpu().numpy(), env.fov.cpu().numpy(), env.adj_matrix.cpu().numpy()
        for e in range(E):
            if not alive[e]:
                continue
            obs[e], d = oracles[e].step(acts[e])
            assert np.array_equal(pos[e], oracles[e].positions), (t, e)
            assert np.array_equal(fov[e], obs[e]["fov"].astype(np.float32)), (t, e)
            assert np.array_equal(adj[e], obs[e]["adj_matrix"].astype(np.float32)), (t, e)
            alive[e] = not d
        assert np.array_equal(env.done.cpu().numpy().astype(bool), ~alive)
    assert flips <= 2


def test_rollout_multi_step_equals_single_steps():
    import mapf_gnn_b200 as m
    E, N, H, M = 40, 20, 32, 40
    rng = np.random.default_rng(2)
    starts, goals, obst = eo.make_scenarios(rng, E, N, H, M)
    cfg = model_config("gnn_msg_k3", N, board_size=[H, H])
    net = load_network("gnn_msg_k3", num_agents=N).eval()
    env_a = m.BatchedGraphEnv(cfg, goals, starts, obst)
    env_b = m.BatchedGraphEnv(cfg, goals, starts, obst)
    ra, rb = m.Rollout(net, env_a), m.Rollout(net, env_b)
    ra.run(8)
    for _ in range(8):
        rb.run(1)
    assert torch.equal(env_a.pos, env_b.pos) and torch.equal(env_a.fov, env_b.fov)
    assert torch.equal(env_a.adj_matrix, env_b.adj_matrix) and torch.equal(env_a.time, env_b.time)


def test_batched_evaluator_matches_oracle_protocol():
    """evaluate.py:201-271 for a whole batch at once: metrics must equal the oracle env driven by the oracle net
    (teacher-forcing is not possible here, so compare only episodes whose every decision had a wide margin)."""
    import mapf_gnn_b200 as m
    cfg = model_config("gnn_k3", 5, board_size=[18, 18], obstacles=5, max_steps=12, tests_episodes=40, sensing_range=6)
    net = load_network("gnn_k3").eval()
    summary, per = m.evaluate(net, cfg, seed=5)
    assert summary["episodes"] == 40 and 0.0 <= summary["avg_success_rate"] <= 1.0
    assert per["steps_taken"].max() <= 22 and per["flow_time"].shape == (40,)
    rng = np.random.default_rng(5)
    starts, goals, obst = m.make_scenarios(rng, 40, 5, 18, 5)
    params = oracle_params("gnn_k3")
    checked = 0
    for e in range(40):
        env = eo.EnvOracle(5, 18, goals[e], starts[e], obst[e], sensing_range=6, max_time=cfg["max_time"])
        obs, safe, done = env.observe(), True, False
        for t in range(22):
            st = torch.tensor(obs["fov"]).float()[None]
            gs = torch.tensor(obs["adj_matrix"]).float()[None]
            with torch.no_grad():
                lg = mo.forward(params, st, gs).numpy()[0]
            safe = safe and top2_margin(lg).min() > 1e-3
            obs, done = env.step(lg.argmax(-1))
            if done:
                break
        if safe:
            s, f = env.metrics()
            assert abs(per["success_rate"][e] - s) < 1e-6 and int(per["flow_time"][e]) == f
            assert int(per["steps_taken"][e]) == env.time and bool(per["completed"][e]) == done
            checked += 1
    assert checked >= 20


def test_evaluator_with_device_scenarios():
    """Same evaluator fed by the on-device scenario generator: deterministic in the seed, plausible summary."""
    import mapf_gnn_b200 as m
    cfg = model_config("gnn_k3", 5, board_size=[18, 18], obstacles=5, max_steps=12, sensing_range=6)
    net = load_network("gnn_k3").eval()
    s1, p1 = m.evaluate(net, cfg, episodes=300, seed=3, device_scenarios=True)
    s2, p2 = m.evaluate(net, cfg, episodes=300, seed=3, device_scenarios=True)
    strip = lambda d: {k: v for k, v in d.items() if k != "batch_time_s"}
    assert strip(s1) == strip(s2) and np.array_equal(p1["flow_time"], p2["flow_time"])
    from mapf_gnn_b200.evaluate import results_dict
    import json
    doc = json.loads(json.dumps(results_dict(cfg, s1, p1)))                  # evaluate.py:401-422 schema
    assert set(doc) == {"timestamp", "config", "summary", "episodes"} and len(doc["episodes"]) == 300
    assert set(doc["episodes"][0]) == {"success_rate", "flow_time", "steps_taken", "inference_time_ms", "episode_time_s",
                                       "collisions", "agents_at_goal", "total_agents"}
    assert set(doc["summary"]) == {"complete_success", "avg_success_rate", "std_success_rate", "avg_flow_time", "avg_steps",
                                   "avg_inference_ms"}
    sv, pv = m.evaluate(net, cfg, episodes=300, seed=3, validation=True)      # train.py:105-141: max_steps, no +10
    assert pv["steps_taken"].max() <= cfg["max_steps"] and 0.0 <= sv["avg_success_rate"] <= 1.0
    s3, _ = m.evaluate(net, cfg, episodes=300, seed=3)
    assert abs(s1["avg_success_rate"] - s3["avg_success_rate"]) < 0.12      # same distribution family


@pytest.mark.parametrize("zero_copy", [False, True])
@pytest.mark.parametrize("E", [64, 2000])
def test_host_stepped_rollout_matches_device_rollout(zero_copy, E):
    """HostSteppedRollout (pinned host buffers + CUDA-graph replay) must walk the same trajectory as Rollout.run --
    with copy nodes around the kernels, and with the kernels reading / writing the pinned block themselves (zero_copy);
    64 episodes take the env step fused into the graph kernel, 2000 the separate env kernel."""
    import mapf_gnn_b200 as m
    N, H, M = 5, 18, 5
    rng = np.random.default_rng(9)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, M)
    cfg = model_config("gnn_k3", N, board_size=[H, H])
    net = load_network("gnn_k3").eval()
    env_a = m.BatchedGraphEnv(cfg, goals, starts, obst)
    env_b = m.BatchedGraphEnv(cfg, goals, starts, obst)
    ra, rb = m.Rollout(net, env_a), m.Rollout(net, env_b)
    hs = m.HostSteppedRollout(rb, zero_copy=zero_copy)
    for t in range(6):
        if t == 3:      # the host edits the state it holds: swap two episodes' positions (on both sides)
            p = hs.pos_in
            p[[0, 1]] = p[[1, 0]].clone()
            env_a.pos[[0, 1]] = env_a.pos[[1, 0]].clone()
        ra.run(1)
        actions, pos_out, done = hs.step()
        assert torch.equal(actions, ra.actions.cpu()) and torch.equal(pos_out, env_a.pos.cpu())
        assert torch.equal(done, env_a.done.cpu())
        assert hs.pos_in.data_ptr() == pos_out.data_ptr()      # the result is, in place, the next step's input
    assert torch.equal(env_a.fov, env_b.fov) and torch.equal(env_a.time, env_b.time)
    hs.reset()                                                 # device state and pinned state back to the scenario
    env_a.reset()
    ra.run(1)
    actions, pos_out, done = hs.step()
    assert torch.equal(actions, ra.actions.cpu()) and torch.equal(pos_out, env_a.pos.cpu())


@pytest.mark.parametrize("path", ["fused_tc", "unfused_tc", "ffma", "tc_fc", "ffma_encoder"])
@pytest.mark.parametrize("name,N", [("gnn_k3", 5), ("gnn_msg_k3", 20)])
def test_every_policy_path_matches_oracle(monkeypatch, path, name, N):
    """The graph-filter / FC stage has several implementations (fused tcgen05 kernel, taps + tcgen05 GEMM, fused FFMA
    kernel, tensor-core FC): each must reproduce the oracle logits."""
    from mapf_gnn_b200 import ops
    if path == "unfused_tc":
        monkeypatch.setenv("MAPF_B200_UNFUSED_GRAPH", "1")
    if path == "ffma":
        monkeypatch.setattr(ops, "USE_TENSOR_CORES", False)
    if path == "tc_fc":            # FFMA2 conv kernel emitting the pre-split operand + tensor-core FC
        monkeypatch.setattr(ops, "USE_TENSOR_CORE_FC", True)
        monkeypatch.setenv("MAPF_B200_ENC_FFMA_CONV", "1")
    if path == "ffma_encoder":     # fused FFMA2 encoder kernel (the default is the tensor-core conv stack)
        monkeypatch.setattr(ops, "USE_TENSOR_CORE_ENCODER", False)
    rng = np.random.default_rng(17 + N)
    B = 70
    states = torch.tensor(rng.integers(0, 4, size=(B, N, 2, 5, 5))).float()
    adj = torch.tensor((rng.random((B, N, N)) < 0.3).astype(np.float32))
    adj[:, torch.arange(N), torch.arange(N)] = 0
    adj[3, 2, :] = 0
    with torch.no_grad():
        ref = mo.forward(oracle_params(name), states, adj).numpy()
    net = load_network(name, num_agents=N).eval()
    logits, actions = net.act(states.cuda(), adj.cuda())
    assert rel_err(logits.cpu().numpy(), ref) < REL
    _check_actions(actions.cpu().numpy(), ref)


@pytest.mark.parametrize("N", [1, 2, 3, 4, 8, 16, 17, 31, 32, 33, 48, 64])
@pytest.mark.parametrize("name", ["gnn_k3", "gnn_msg_k3"])
def test_team_sizes_sweep(name, N):
    """Every tiling branch of the graph stage (tile of 32 / 64 / 128 rows, fused vs two-kernel tensor-core path,
    4-aligned vs generic hops) against the oracle, including N = 1 (no neighbours) and ragged last tiles."""
    rng = np.random.default_rng(100 + N)
    B = 37 if N <= 16 else 9
    states = torch.tensor(rng.integers(0, 4, size=(B, N, 2, 5, 5))).float()
    adj = torch.tensor((rng.random((B, N, N)) < min(0.5, 4.0 / max(N, 1))).astype(np.float32))
    adj[:, torch.arange(N), torch.arange(N)] = 0
    with torch.no_grad():
        ref = mo.forward(oracle_params(name), states, adj).numpy()
    net = load_network(name, num_agents=N).eval()
    logits, actions = net.act(states.cuda(), adj.cuda())
    assert rel_err(logits.cpu().numpy(), ref) < REL
    _check_actions(actions.cpu().numpy(), ref)


@pytest.mark.parametrize("E,N,H", [(1, 1, 3), (1, 5, 18), (3, 2, 4), (129, 3, 6), (17, 128, 40)])
def test_env_edge_sizes(E, N, H):
    """Single episode, single agent, episode counts that do not fill a CTA, the largest supported team."""
    import mapf_gnn_b200 as m
    rng = np.random.default_rng(E * 7 + N)
    M = min(3, H * H - 2 * N) if H * H > 2 * N + 3 else 0
    starts, goals, obst = m.make_scenarios(rng, E, N, H, M)
    cfg = {"num_agents": N, "board_size": [H, H], "max_time": 16, "min_time": 1}
    env = m.BatchedGraphEnv(cfg, goals, starts, obst if M else None, sensing_range=3)
    oracles = [eo.EnvOracle(N, H, goals[e], starts[e], obst[e] if M else None, sensing_range=3, max_time=16)
               for e in range(E)]
    alive = np.ones(E, dtype=bool)
    for t in range(5):
        actions = rng.integers(0, 5, size=(E, N))
        obs, done = env.step(actions)
        fov, adj, pos = obs["fov"].cpu().numpy(), obs["adj_matrix"].cpu().numpy(), env.pos.cpu().numpy()
        for e in range(E):
            if not alive[e]:
                continue
            o, d = oracles[e].step(actions[e])
            assert np.array_equal(pos[e], oracles[e].positions)
            assert np.array_equal(fov[e], o["fov"].astype(np.float32))
            assert np.array_equal(adj[e], o["adj_matrix"].astype(np.float32))
            alive[e] = not d


@pytest.mark.parametrize("zero_copy", [False, True])
def test_pipelined_host_rollout_matches_device_rollout(zero_copy):
    """PipelinedHostRollout: groups stepped with overlapping host round trips walk the same trajectories as one
    on-device Rollout over all episodes (episodes are independent)."""
    import mapf_gnn_b200 as m
    E, N, H, M = 90, 5, 18, 5
    rng = np.random.default_rng(21)
    starts, goals, obst = m.make_scenarios(rng, E, N, H, M)
    cfg = model_config("gnn_k3", N, board_size=[H, H])
    net = load_network("gnn_k3").eval()
    env = m.BatchedGraphEnv(cfg, goals, starts, obst)
    ro = m.Rollout(net, env)
    assert m.PipelinedHostRollout.round_aligned_sizes(4096, 5, 2, 148) == [1894, 2202]
    assert m.PipelinedHostRollout.round_aligned_sizes(90, 5, 3, 148) == [30, 30, 30]
    pl = m.PipelinedHostRollout(net, cfg, goals, starts, obst, groups=3, zero_copy=zero_copy)
    assert [s.stop - s.start for s in pl.slices] == [30, 30, 30]
    pl.start()
    for t in range(5):
        ro.run(1)
        for g, sl in enumerate(pl.slices):
            actions, pos_out, done = pl.advance(g, relaunch=t < 4)
            assert torch.equal(actions, ro.actions.cpu()[sl]) and torch.equal(pos_out, env.pos.cpu()[sl])
            assert torch.equal(done, env.done.cpu()[sl])
    assert pl.drain() == [None, None, None]
    pl.reset()
    assert torch.equal(pl.envs[1].pos.cpu(), torch.as_tensor(starts[30:60]).int())
    with pytest.raises(RuntimeError):
        pl.groups[0].wait()


@pytest.mark.parametrize("msg_type", ["gcn", "message"])
def test_general_configs_staged_inference_matches_oracle(msg_type):
    """framework_gnn.py:90-123 allows deeper encoder MLPs (`encoder_layers`) and several graph filtering layers
    (`graph_filters` / `node_dims` lists): no shipped config uses them, they run stage by stage on the same kernels
    (inference only).  Logits vs the torch-CPU oracle on the module's own state_dict; action_layers > 1 cannot be built
    by the reference either (IndexError) and raises the same here."""
    import mapf_gnn_b200 as m
    torch.manual_seed(5)
    N, B = 7, 33
    cfg = dict(map_shape=[5, 5], num_actions=5, last_convs=[400], channels=[16, 16, 16], node_dims=[128, 64],
               graph_filters=[3, 2], encoder_dims=[64, 32], encoder_layers=2, action_layers=1, num_agents=N,
               net_type="gnn", msg_type=msg_type)
    net = m.build_network(cfg).cuda().eval()
    assert not net.fused
    for mod in net.modules():                       # non-trivial running statistics
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0.0, 0.3)
            mod.running_var.uniform_(0.5, 1.5)
    params = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    assert {"compressMLP.2.weight", "GFL.2.W"} <= set(params)
    rng = np.random.default_rng(8)
    states = torch.tensor(rng.integers(0, 4, size=(B, N, 2, 5, 5))).float()
    adj = torch.tensor((rng.random((B, N, N)) < 0.35).astype(np.float32))
    adj[:, torch.arange(N), torch.arange(N)] = 0
    with torch.no_grad():
        ref = mo.forward(params, states, adj).numpy()
        out = net(states.cuda(), adj.cuda())
        logits, actions = net.act(states.cuda(), adj.cuda())
    assert rel_err(out.cpu().numpy(), ref) < REL
    assert torch.equal(out, logits)
    assert _check_actions(actions.cpu().numpy(), ref) == 0
    with pytest.raises(NotImplementedError):
        net.train()(states.cuda(), adj.cuda())
    with pytest.raises(IndexError):
        m.build_network(dict(cfg, action_layers=2))
