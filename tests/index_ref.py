"""Independent plain-Python restatement of the nnet3 index regularisation used by
TdnnDARTSV3Component (tdnn.cc:822-905).  Test infrastructure only."""
from math import gcd

K_NO_TIME = -32768


def get_computation_io(inp, out):
    nx = sorted({(n, x) for (n, t, x) in inp})
    t_in = sorted({t for (_, t, _) in inp if t != K_NO_TIME})
    t_out = sorted({t for (_, t, _) in out if t != K_NO_TIME})

    def regularize(ts):
        g = 0
        for a, b in zip(ts, ts[1:]):
            g = gcd(g, b - a)
        return (g, len(ts)) if g == 0 else (g, 1 + (ts[-1] - ts[0]) // g)

    step_in, num_in = regularize(t_in)
    step_out, num_out = regularize(t_out)
    return dict(num_images=len(nx), nx=nx, start_t_in=t_in[0], t_step_in=step_in, num_t_in=num_in,
                start_t_out=t_out[0], t_step_out=step_out, num_t_out=num_out, reorder_t_in=1)


def modify_computation_io(io):
    if io["t_step_out"] == 0:
        if io["t_step_in"] == 0:
            io["t_step_in"] = 1
        io["t_step_out"] = io["t_step_in"]
    assert io["t_step_out"] % io["t_step_in"] == 0
    r = io["t_step_out"] // io["t_step_in"]
    io["reorder_t_in"] = r
    io["num_t_in"] = r * ((io["num_t_in"] + r - 1) // r)
    return io


def _create(nx, start, step, num, reorder):
    if step == 0:
        return [(n, start, x) for (n, x) in nx]
    out = []
    for b in range(num // reorder):
        for (n, x) in nx:
            for k in range(reorder):
                out.append((n, start + (b * reorder + k) * step, x))
    return out


def get_indexes_for_computation(io, inp, out):
    si, so = set(inp), set(out)
    new_in = [i if i in si else (i[0], K_NO_TIME, i[2])
              for i in _create(io["nx"], io["start_t_in"], io["t_step_in"], io["num_t_in"], io["reorder_t_in"])]
    new_out = [i if i in so else (i[0], K_NO_TIME, i[2])
               for i in _create(io["nx"], io["start_t_out"], io["t_step_out"], io["num_t_out"], 1)]
    return new_in, new_out


def reorder_indexes(inp, out):
    io = modify_computation_io(get_computation_io(inp, out))
    return get_indexes_for_computation(io, inp, out)


def precompute_indexes(time_offsets, inp, out):
    io = modify_computation_io(get_computation_io(inp, out))
    r = io["reorder_t_in"]
    offs = []
    for off in time_offsets:
        req = io["start_t_out"] + off
        input_t = (req - io["start_t_in"]) // io["t_step_in"]
        assert req == io["start_t_in"] + io["t_step_in"] * input_t
        offs.append(r * (input_t // r) * io["num_images"] + input_t % r)
    return r, offs
