"""Config 4, host side: beam width sweep 1/10/50/200 on T=4000 utterances with the CPU restatement of pyctcdecode
(oracle/beam.py; the library itself is unavailable) over all host cores, one utterance per process at a time.
N = one utterance per core (N is printed); the GPU side of the same sweep is in profiles/sweeps.py.
    python profiles/cfg4_cpu.py > profiles/r01_cfg4_cpu.json"""
import json
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import beam as ob
from silent_speech_decoder_b200 import synth
from silent_speech_decoder_b200.data.vocab import Vocab


def work(job):
    return ob.decode_one(job)


def main():
    v = Vocab.default()
    cores = os.cpu_count() or 1
    T = 4000
    out = []
    for kind in ("peaked", "flat"):
        n = cores if kind == "peaked" else max(4, cores // 2)
        lp = (synth.peaked(n, T, 1004) if kind == "peaked" else synth.flat(n, T, 1004)).numpy()
        for w in (1, 10, 50, 200):
            dec = ob.BeamDecoder(v.tokens, v.blank_id, v.pad_id, w, -10.0)
            mat = ob.wrapper_matrix(lp, v.blank_id, v.pad_id, 0.0, dec.keep)
            jobs = [(mat[i], dec.labels, w, -10.0, -5.0) for i in range(n)]
            with mp.get_context("fork").Pool(cores) as pool:
                t0 = time.perf_counter()
                pool.map(work, jobs, chunksize=1)
                dt = time.perf_counter() - t0
            out.append({"input": kind, "N": n, "T": T, "width": w, "seconds": dt, "utt_per_s": n / dt, "cores": cores,
                        "kind": "port (oracle/beam.py)"})
            print(out[-1], file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
