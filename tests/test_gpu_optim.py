"""ArenaAdam (one fused launch over the flat parameter arena) against torch.optim.Adam on the CPU - the optimiser the
reference constructs at train_MuRCL.py:154-171 - over several steps, with weight decay, incl. the bf16 shadow weights and
CUDA-graph replay (device-side step counter)."""
import pytest
import torch

from murcl_b200.arena import ParamArena
from murcl_b200.optim import ArenaAdam
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _modules(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(96, 130), torch.nn.ReLU(), torch.nn.Linear(130, 7), torch.nn.GRU(7, 33))


def _grads(step, params):
    g = torch.Generator().manual_seed(100 + step)
    return [torch.randn(p.shape, generator=g) * (0.5 + step) for p in params]


@pytest.mark.parametrize("wd", [0.0, 1e-5, 0.1])
def test_arena_adam_matches_torch_adam(wd):
    ref = _modules(3)
    dut = _modules(3).to(DEV)
    arena = ParamArena(list(dut.parameters()))
    opt = ArenaAdam(arena, lr=3e-3, weight_decay=wd)
    ref_opt = torch.optim.Adam(ref.parameters(), lr=3e-3, weight_decay=wd)
    for step in range(6):
        gs = _grads(step, list(ref.parameters()))
        opt.zero_grad()
        for p, q, g in zip(ref.parameters(), dut.parameters(), gs):
            p.grad = g.clone()
            q.grad.add_(g.to(DEV))                      # kernels ADD into the arena's gradient views
        ref_opt.step()
        opt.step()
    assert opt.steps == 6
    for (n, p), q in zip(ref.named_parameters(), dut.parameters()):
        assert_close(q, p, 2e-6, f"param {n} after 6 steps (wd={wd})")
        sh = q._murcl_shadow
        assert sh.dtype == torch.bfloat16
        assert torch.equal(sh, q.detach().to(torch.bfloat16)), f"shadow of {n} is not the rounded parameter"
    st = ref_opt.state[next(iter(ref.parameters()))]
    n0 = next(iter(ref.parameters())).numel()
    assert_close(opt.exp_avg[:n0].view_as(st["exp_avg"]), st["exp_avg"], 2e-6, "exp_avg")
    assert_close(opt.exp_avg_sq[:n0].view_as(st["exp_avg_sq"]), st["exp_avg_sq"], 2e-6, "exp_avg_sq")


def test_arena_adam_replays_in_a_cuda_graph():
    """The step counter is device state: N replays of one captured step == N eager steps (bias correction advances)."""
    a, b = _modules(5).to(DEV), _modules(5).to(DEV)
    arena_a, arena_b = ParamArena(list(a.parameters())), ParamArena(list(b.parameters()))
    opt_a, opt_b = ArenaAdam(arena_a, lr=1e-2, weight_decay=1e-4), ArenaAdam(arena_b, lr=1e-2, weight_decay=1e-4)
    g = torch.randn_like(arena_a.grad)
    arena_a.grad.copy_(g); arena_b.grad.copy_(g)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        opt_a.step()
    for _ in range(4):
        graph.replay()
    for _ in range(4):
        opt_b.step()
    torch.cuda.synchronize()
    assert opt_a.steps == 4 and opt_b.steps == 4
    assert torch.equal(arena_a.flat, arena_b.flat)
    sd = opt_a.state_dict()
    opt_b.load_state_dict(sd)
    assert opt_b.steps == 4


def test_arena_adam_rejects_other_params():
    with pytest.raises(TypeError):
        ArenaAdam([torch.zeros(4, device=DEV)])
