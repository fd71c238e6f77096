"""The library never allocates, never synchronises and only enqueues on the stream it is given, so the
operators (forward and backward) can be captured in a CUDA graph and replayed on new data."""
import pytest
import torch

from relation_detr_b200 import ops, workloads

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_msda_and_relation_ops_capture_and_replay_in_a_cuda_graph():
    shape = workloads.MsdaShape("t", 2, ((25, 42), (13, 21), (7, 11), (4, 6)), 300)
    a = workloads.make_msda_inputs(shape, "S", seed=0, device=DEV)
    b = workloads.make_msda_inputs(shape, "oob", seed=1, device=DEV)
    r1 = workloads.make_rel_inputs(workloads.RelShape("t", 2, 70, 45), seed=0, device=DEV)
    r2 = workloads.make_rel_inputs(workloads.RelShape("t", 2, 70, 45), seed=1, device=DEV)
    dim_t = ops.relation_dim_t(16, 10000.0, DEV)
    static = {k: v.clone() for k, v in a.items()}
    srel = {k: v.clone() for k, v in r1.items()}

    def run():
        out = ops.msda_forward(static["value"], static["spatial_shapes"], static["level_start_index"],
                               static["sampling_locations"], static["attention_weights"])
        gv, gl, ga = ops.msda_backward(static["value"], static["spatial_shapes"], static["level_start_index"],
                                       static["sampling_locations"], static["attention_weights"], static["grad_output"])
        bias, bits = ops.relation_forward(srel["src_boxes"], srel["tgt_boxes"], srel["weight"], srel["bias"], dim_t, 100.0, 1e-5, None, True)
        gw, gb = ops.relation_backward(srel["src_boxes"], srel["tgt_boxes"], dim_t, 100.0, 1e-5, srel["grad_output"], bits, 8, True)
        return out, gv, gl, ga, bias, gw, gb

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up on the capture stream (lazy module load, attribute calls)
        run()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        captured = run()
    for data, rel in ((a, r1), (b, r2)):
        for k in static:
            static[k].copy_(data[k])
        for k in srel:
            srel[k].copy_(rel[k])
        graph.replay()
        torch.cuda.synchronize()
        got = [t.clone() for t in captured]
        want = run()
        torch.cuda.synchronize()
        assert torch.equal(got[0], want[0])                      # out
        assert (got[1] - want[1]).abs().max().item() <= 1e-5     # grad_value: atomic order differs run to run
        assert torch.equal(got[2], want[2]) and torch.equal(got[3], want[3])
        assert torch.equal(got[4], want[4])                      # relation bias
        assert (got[5] - want[5]).abs().max().item() <= 1e-3 * max(1.0, want[5].abs().max().item())
