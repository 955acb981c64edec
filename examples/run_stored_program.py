"""Run a stored Blackbird (.xbb) or XIR (.xir) program on the b200fock backend without Strawberry Fields'
`blackbird` / `xir` dependencies (strawberryfields_b200.io).

    python examples/run_stored_program.py program.xbb --cutoff 5 [--ir blackbird|xir] [--save-state ckpt.npz]
                                            [--arg r=0.3 --arg theta=0.4]      # values of template parameters {r}, {theta}

The reference equivalent is `prog = sf.load("program.xbb"); sf.Engine("fock", backend_options={"cutoff_dim": 5}).run(prog)`
(strawberryfields/io/__init__.py:169).  Needs a CUDA device: there is no CPU fallback.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from strawberryfields_b200 import B200FockBackend, io

    ap = argparse.ArgumentParser()
    ap.add_argument("program")
    ap.add_argument("--cutoff", type=int, default=None, help="cutoff_dim (default: the script's own option, if any)")
    ap.add_argument("--ir", default=None, choices=["blackbird", "xir"])
    ap.add_argument("--save-state", default=None, help="write the final state to this .npz checkpoint")
    ap.add_argument("--arg", action="append", default=[], metavar="NAME=VALUE",
                    help="value of a template parameter {NAME} (engine.run(prog, args={..}) in the reference)")
    args = ap.parse_args()
    values = {k: complex(v) if "j" in v else float(v) for k, v in (item.split("=", 1) for item in args.arg)}
    ir = args.ir or ("xir" if args.program.endswith(".xir") else "blackbird")
    prog = io.load(args.program, ir=ir)
    print("program %r: %d modes, %d operations%s%s" % (
        prog.name, prog.num_subsystems, len(prog.operations),
        ", free parameters %s" % prog.free_parameters if prog.is_template else "",
        ", measured parameters fed forward" if prog.has_feed_forward else ""))
    backend = B200FockBackend()
    samples = prog.run(backend, cutoff_dim=args.cutoff, args=values)
    for mode in sorted(samples):
        print("  mode %d: %s" % (mode, samples[mode]))
    state = backend.state()
    print("trace %.12f, pure %s, mean photon numbers %s" % (
        state.trace(), state.is_pure, [round(float(state.mean_photon(m)[0]), 6) for m in range(prog.num_subsystems)]))
    if args.save_state:
        io.save_state(args.save_state, state)
        print("state written to", args.save_state)


if __name__ == "__main__":
    main()
