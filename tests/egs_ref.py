"""Independent Python WRITER of the Kaldi / OpenFst on-disk formats that csrc/egs_io.cc reads (test infrastructure).

Written from the published formats (kaldi: base/io-funcs-inl.h, matrix/{kaldi-matrix,compressed-matrix,sparse-matrix}.cc,
nnet3/{nnet-common,nnet-example,nnet-chain-example}.cc, chain/chain-supervision.cc, fstext/kaldi-fst-io.cc; OpenFst:
fst.h FstHeader, vector-fst.h, compact-fst.h).  Kaldi is not available here, so reader and writer pin each other, not
Kaldi: a file Kaldi itself wrote has never been through this code.
"""
from __future__ import annotations

import struct

import numpy as np

FST_MAGIC = 2125659606


def tok(s: str) -> bytes:
    return s.encode() + b" "


def i32(v: int, binary: bool) -> bytes:
    return b"\x04" + struct.pack("<i", int(v)) if binary else f"{int(v)} ".encode()


def f32(v: float, binary: bool) -> bytes:
    return b"\x04" + struct.pack("<f", float(v)) if binary else (repr(float(np.float32(v))) + " ").encode()


def boolean(v: bool, binary: bool) -> bytes:
    return (b"T" if v else b"F") + (b"" if binary else b" ")


def fnum(x) -> str:
    return repr(float(np.float32(x)))


# ---- vectors / matrices
def float_vector(v, binary: bool) -> bytes:
    v = np.asarray(v, np.float32)
    if binary:
        return tok("FV") + i32(len(v), True) + v.tobytes()
    return (" [ " + " ".join(fnum(x) for x in v) + (" " if len(v) else "") + "]\n").encode()


def integer_vector(v, binary: bool, dtype=np.int32) -> bytes:
    v = np.asarray(v, dtype)
    if binary:
        return bytes([v.dtype.itemsize]) + struct.pack("<i", len(v)) + v.tobytes()
    return ("[ " + " ".join(str(int(x)) for x in v) + (" " if len(v) else "") + "]\n").encode()


def full_matrix(m, binary: bool, double: bool = False) -> bytes:
    m = np.asarray(m, np.float64 if double else np.float32)
    if binary:
        return tok("DM" if double else "FM") + i32(m.shape[0], True) + i32(m.shape[1], True) + np.ascontiguousarray(m).tobytes()
    if m.size == 0:
        return b" [ ]\n"
    return (" [" + "".join("\n  " + " ".join(fnum(x) for x in row) + " " for row in m) + "]\n").encode()


def _u16(h_min, h_range, x):
    f = (np.asarray(x, np.float64) - h_min) / h_range if h_range > 0 else np.zeros_like(np.asarray(x, np.float64))
    return np.clip(np.floor(f * 65535 + 0.499), 0, 65535).astype(np.uint16)


def compressed_matrix(m, fmt: int) -> bytes:
    """CompressedMatrix binary (matrix/compressed-matrix.cc): 1 = one byte + per-column quartile headers ("CM"),
    2 = two bytes ("CM2"), 3 = one byte ("CM3").  Returns the bytes; decode_compressed gives the values a reader must see."""
    m = np.asarray(m, np.float32)
    R, Cc = m.shape
    lo, hi = (float(m.min()), float(m.max())) if m.size else (0.0, 0.0)
    rng = hi - lo if hi > lo else 1.0e-05
    head = struct.pack("<ffii", lo, rng, R, Cc)
    if fmt == 1:
        cols = bytearray()
        data = bytearray()
        for j in range(Cc):
            col = np.sort(m[:, j].astype(np.float64))
            q = [col[0], col[R // 4], col[(3 * R) // 4], col[R - 1]]
            u = _u16(lo, rng, q).astype(np.int64)
            # strictly increasing quartile codes, as Kaldi enforces, so that every piece has a slope
            u[0] = min(u[0], 65532)
            u[1] = min(max(u[1], u[0] + 1), 65533)
            u[2] = min(max(u[2], u[1] + 1), 65534)
            u[3] = max(u[3], u[2] + 1)
            cols += struct.pack("<4H", *[int(x) for x in u])
            p = lo + rng * 1.52590218966964e-05 * u.astype(np.float64)
            x = m[:, j].astype(np.float64)
            b = np.where(x < p[1], np.rint((x - p[0]) / (p[1] - p[0]) * 64),
                         np.where(x < p[2], 64 + np.rint((x - p[1]) / (p[2] - p[1]) * 128), 192 + np.rint((x - p[2]) / (p[3] - p[2]) * 63)))
            data += np.clip(b, 0, 255).astype(np.uint8).tobytes()  # column-major
        return tok("CM") + head + bytes(cols) + bytes(data)
    if fmt == 2:
        return tok("CM2") + head + _u16(lo, rng, m).tobytes()
    b = np.clip(np.rint((m.astype(np.float64) - lo) / rng * 255), 0, 255).astype(np.uint8)
    return tok("CM3") + head + b.tobytes()


def decode_compressed(blob: bytes) -> np.ndarray:
    """What a reader must produce for compressed_matrix(...)'s bytes, in float32 arithmetic as Kaldi's CharToFloat /
    Uint16ToFloat do it."""
    sp = blob.index(b" ")
    token, body = blob[:sp].decode(), blob[sp + 1:]
    lo, rng, R, Cc = struct.unpack("<ffii", body[:16])
    lo, rng = np.float32(lo), np.float32(rng)
    body = body[16:]
    if token == "CM":
        q = np.frombuffer(body[:8 * Cc], np.uint16).reshape(Cc, 4)
        p = (lo + rng * np.float32(1.52590218966964e-05) * q.astype(np.float32)).astype(np.float32)  # [C, 4]
        b = np.frombuffer(body[8 * Cc:8 * Cc + R * Cc], np.uint8).reshape(Cc, R).T.astype(np.float32)  # [R, C]
        p0, p1, p2, p3 = (p[:, k][None, :] for k in range(4))
        f = np.float32
        return np.where(b <= 64, p0 + (p1 - p0) * b * f(1 / 64.0),
                        np.where(b <= 192, p1 + (p2 - p1) * (b - 64) * f(1 / 128.0), p2 + (p3 - p2) * (b - 192) * f(1 / 63.0))).astype(np.float32)
    if token == "CM2":
        u = np.frombuffer(body[:2 * R * Cc], np.uint16).reshape(R, Cc).astype(np.float32)
        return (lo + u * (rng * np.float32(1.0 / 65535.0))).astype(np.float32)
    u = np.frombuffer(body[:R * Cc], np.uint8).reshape(R, Cc).astype(np.float32)
    return (lo + u * (rng * np.float32(1.0 / 255.0))).astype(np.float32)


def sparse_matrix(m, binary: bool) -> bytes:
    m = np.asarray(m, np.float32)
    out = (tok("SM") + i32(m.shape[0], True)) if binary else f"rows={m.shape[0]} ".encode()
    for row in m:
        nz = np.nonzero(row)[0]
        if binary:
            out += tok("SV") + i32(m.shape[1], True) + i32(len(nz), True) + b"".join(i32(k, True) + f32(row[k], True) for k in nz)
        else:
            out += (f"dim={m.shape[1]} [ " + "".join(f"{k} {fnum(row[k])} " for k in nz) + "] ").encode()
    return out


# ---- nnet3 index vectors (nnet-common.cc WriteIndexVector)
def index_vector(indexes, binary: bool) -> bytes:
    indexes = [tuple(int(v) for v in ix) for ix in indexes]
    out = tok("<I1V>") + i32(len(indexes), binary)
    for k, (n, t, x) in enumerate(indexes):
        if not binary:
            out += tok("<I1>") + i32(n, False) + i32(t, False) + i32(x, False)
            continue
        if k == 0:
            small = n == 0 and x == 0 and abs(t) < 125
            step = t
        else:
            pn, pt, px = indexes[k - 1]
            small = n == pn and x == px and abs(t - pt) < 125
            step = t - pt
        out += struct.pack("<b", step) if small else (struct.pack("<b", 127) + i32(n, True) + i32(t, True) + i32(x, True))
    return out


# ---- FSTs: dict(start, num_states, arcs=[(src, dst, ilabel, weight)], finals={state: weight})
def fst_text_lines(fst) -> str:
    """fstprint: arcs grouped by state, the start state's lines first, a final state's line after its arcs."""
    order = [fst["start"]] + [s for s in range(fst["num_states"]) if s != fst["start"]]
    lines = []
    for s in order:
        for (a, b, il, w) in fst["arcs"]:
            if a == s:
                lines.append(f"{a}\t{b}\t{il}\t{il}" + ("" if w == 0 else f"\t{fnum(w)}"))
        if s in fst["finals"]:
            w = fst["finals"][s]
            lines.append(f"{s}" + ("" if w == 0 else f"\t{fnum(w)}"))
    return "\n".join(lines) + "\n"


def fst_kaldi_text(fst) -> bytes:
    """WriteFstKaldi, text mode: a newline, the fstprint lines, an empty line."""
    return ("\n" + fst_text_lines(fst) + "\n").encode()


def _fst_header(fsttype: str, start: int, nstates: int, narcs: int, version: int = 2, flags: int = 0) -> bytes:
    s = lambda x: struct.pack("<i", len(x)) + x.encode()
    return (struct.pack("<i", FST_MAGIC) + s(fsttype) + s("standard") + struct.pack("<iiQqqq", version, flags, 0, start, nstates, narcs))


def fst_compact_acceptor(fst) -> bytes:
    """CompactFst<StdArc, AcceptorCompactor>::Write: header, uint32 offsets [nstates + 1], elements (label, weight, nextstate);
    a final state's first element is (-1, final weight, -1)."""
    offs, elems = [0], b""
    n = 0
    for s in range(fst["num_states"]):
        if s in fst["finals"]:
            elems += struct.pack("<ifi", -1, fst["finals"][s], -1)
            n += 1
        for (a, b, il, w) in fst["arcs"]:
            if a == s:
                elems += struct.pack("<ifi", il, w, b)
                n += 1
        offs.append(n)
    return _fst_header("compact_acceptor", fst["start"], fst["num_states"], len(fst["arcs"])) + struct.pack(f"<{len(offs)}I", *offs) + elems


def fst_vector(fst) -> bytes:
    """VectorFst<StdArc>::Write: per state the final weight (+inf = not final), int64 arc count, arcs (ilabel, olabel, weight, next)."""
    out = _fst_header("vector", fst["start"], fst["num_states"], len(fst["arcs"]))
    for s in range(fst["num_states"]):
        arcs = [(a, b, il, w) for (a, b, il, w) in fst["arcs"] if a == s]
        out += struct.pack("<fq", fst["finals"].get(s, float("inf")), len(arcs))
        for (_, b, il, w) in arcs:
            out += struct.pack("<iifi", il, il, w, b)
    return out


# ---- examples
def supervision(sup, binary: bool) -> bytes:
    out = (tok("<Supervision>") + tok("<Weight>") + f32(sup["weight"], binary) + tok("<NumSequences>") + i32(sup["num_sequences"], binary)
           + tok("<FramesPerSeq>") + i32(sup["frames_per_seq"], binary) + tok("<LabelDim>") + i32(sup["label_dim"], binary))
    if sup.get("write_e2e_flag", True):
        out += tok("<End2End>") + boolean(sup["e2e"], binary)
    wr = fst_compact_acceptor if binary else fst_kaldi_text
    if not sup["e2e"]:
        out += wr(sup["fsts"][0])
    else:
        out += tok("<Fsts>") + b"".join(wr(f) for f in sup["fsts"]) + tok("</Fsts>")
    if sup.get("write_alignment_pdfs", True):
        out += tok("<AlignmentPdfs>") + integer_vector(sup.get("alignment_pdfs", []), binary)
    return out + tok("</Supervision>")


def general_matrix(m, binary: bool, coding: str) -> bytes:
    if coding == "full" or (not binary and coding.startswith("cm")):
        return full_matrix(m, binary)        # a compressed matrix is written as a plain one in text mode
    if coding == "double":
        return full_matrix(m, binary, double=True)
    if coding == "sparse":
        return sparse_matrix(m, binary)
    return compressed_matrix(m, int(coding[2:]))


def example(ex, binary: bool) -> bytes:
    nl = b"" if binary else b"\n"
    out = tok("<Nnet3ChainEg>") + tok("<NumInputs>") + i32(len(ex["inputs"]), binary) + nl
    for io in ex["inputs"]:
        out += (tok("<NnetIo>") + tok(io["name"]) + index_vector(io["indexes"], binary)
                + general_matrix(io["data"], binary, io.get("coding", "full")) + tok("</NnetIo>") + nl)
    out += tok("<NumOutputs>") + i32(len(ex["outputs"]), binary) + nl
    for sup in ex["outputs"]:
        out += tok("<NnetChainSup>") + tok(sup["name"]) + index_vector(sup["indexes"], binary) + supervision(sup, binary)
        dw = np.asarray(sup.get("deriv_weights", []), np.float32)
        if sup.get("dw2", False):
            out += tok("<DW2>") + float_vector(dw, binary)
        elif binary:   # WriteVectorAsChar
            out += tok("<DW>") + integer_vector(np.floor(255.0 * dw + 0.5), True, np.uint8)
        else:
            out += tok("<DW>") + float_vector(dw, False)
        out += tok("</NnetChainSup>") + nl
    return out + tok("</Nnet3ChainEg>")


def ark(examples, binary: bool) -> bytes:
    return b"".join(tok(ex["key"]) + (b"\0B" if binary else b"") + example(ex, binary) + (b"" if binary else b"\n") for ex in examples)


# ---- synthetic examples shaped like the recipes' (`--constrained false`: one FST per sequence)
def random_fst(rng, frames: int, num_pdfs: int, extra: int = 2) -> dict:
    """A left-to-right acceptor over `frames` frames with a few alternative arcs and self-loop-like skips; labels pdf-id + 1."""
    arcs = []
    for t in range(frames):
        for _ in range(1 + int(rng.integers(0, extra + 1))):
            arcs.append((t, t + 1, int(rng.integers(1, num_pdfs + 1)), float(np.float32(rng.uniform(0, 2))) if rng.random() < 0.5 else 0.0))
    return dict(start=0, num_states=frames + 1, arcs=arcs, finals={frames: 0.0 if rng.random() < 0.5 else float(np.float32(rng.uniform(0, 1)))})


def random_example(rng, key: str, *, frames: int = 6, left: int = 3, right: int = 2, dim: int = 5, ivector_dim: int = 4, num_pdfs: int = 11,
                   e2e: bool = True, num_sequences: int = 1, coding: str = "full", frame_subsampling: int = 3, deriv_weights: bool = True,
                   dw2: bool = False) -> dict:
    t_in = list(range(-left, frames * frame_subsampling + right))
    in_idx = [(n, t, 0) for n in range(num_sequences) for t in t_in]
    inputs = [dict(name="input", indexes=in_idx, data=rng.standard_normal((len(in_idx), dim)).astype(np.float32), coding=coding)]
    if ivector_dim:
        inputs.append(dict(name="ivector", indexes=[(n, 0, 0) for n in range(num_sequences)],
                           data=rng.standard_normal((num_sequences, ivector_dim)).astype(np.float32)))
    out_idx = [(n, t * frame_subsampling, 0) for t in range(frames) for n in range(num_sequences)]
    nf = num_sequences if e2e else 1
    sup = dict(name="output", indexes=out_idx, weight=1.0, num_sequences=num_sequences, frames_per_seq=frames, label_dim=num_pdfs,
               e2e=e2e, fsts=[random_fst(rng, frames, num_pdfs) for _ in range(nf)], alignment_pdfs=[], dw2=dw2,
               deriv_weights=(np.round(rng.uniform(0, 1, frames * num_sequences) * 255) / 255).astype(np.float32) if deriv_weights else [])
    return dict(key=key, inputs=inputs, outputs=[sup])


def den_graph_to_fst(graph, text: str) -> dict:
    """The FST dict of this module from the FSM text synth.den_graph_to_fst_text(graph) writes."""
    fst = dict(start=None, num_states=int(graph["num_states"]), arcs=[], finals={})
    for line in text.splitlines():
        f = line.split()
        if len(f) >= 4:
            fst["arcs"].append((int(f[0]), int(f[1]), int(f[2]), float(f[4]) if len(f) > 4 else 0.0))
            if fst["start"] is None:
                fst["start"] = int(f[0])
        elif f:
            fst["finals"][int(f[0])] = float(f[1]) if len(f) > 1 else 0.0
    return fst
