"""`lzfoo` on the GPU engine: python -m lzfse_rust_b200 -encode|-decode [-i INPUT] [-o OUTPUT] [-v]

The same command line and exit codes as the reference's lzfoo (lzfoo/main.rs:29-109): stdin / stdout when -i / -o are
absent, -v prints sizes, ratio and speed to stderr, exit code 1 with `Error: ...` on a bad frame or an IO error."""
import sys
import time


def _stats(start, n_in, n_out, inp, outp, mode):
    secs = max(time.perf_counter() - start, 1e-12)
    n_raw, n_payload = (n_in, n_out) if mode == "encode" else (n_out, n_in)
    if outp == "stdout":
        print(file=sys.stderr)
    print("LZFSE %s" % mode, file=sys.stderr)
    print("Input: %s" % inp, file=sys.stderr)
    print("Output: %s" % outp, file=sys.stderr)
    print("Input size: %d B" % n_in, file=sys.stderr)
    print("Output size: %d B" % n_out, file=sys.stderr)
    print("Compression ratio: %.3f" % (n_raw / max(n_payload, 1)), file=sys.stderr)
    print("Speed: %.2f ns/B, %.2f MB/s" % (1e9 * secs / max(n_raw, 1), n_raw / secs / 1024.0 / 1024.0), file=sys.stderr)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in ("-encode", "-decode"):
        print("usage: python -m lzfse_rust_b200 -encode|-decode [-i INPUT] [-o OUTPUT] [-v]", file=sys.stderr)
        return 2
    mode, rest = argv[0][1:], argv[1:]
    inp = outp = None
    verbose = False
    i = 0
    while i < len(rest):
        if rest[i] == "-i" and i + 1 < len(rest):
            inp = rest[i + 1]; i += 2
        elif rest[i] == "-o" and i + 1 < len(rest):
            outp = rest[i + 1]; i += 2
        elif rest[i] == "-v":
            verbose = True; i += 1
        else:
            print("Error: unexpected argument %r" % rest[i], file=sys.stderr)
            return 2
    from .streaming import LzfseError, LzfseRingDecoder, LzfseRingEncoder

    try:
        src = open(inp, "rb") if inp else sys.stdin.buffer
        dst = open(outp, "wb") if outp else sys.stdout.buffer
        start = time.perf_counter()
        if mode == "encode":
            n_raw, n_payload = LzfseRingEncoder().encode(src, dst)
            n_in, n_out = n_raw, n_payload
        else:
            n_raw, n_payload = LzfseRingDecoder().decode(src, dst)
            n_in, n_out = n_payload, n_raw
        dst.flush()
        if verbose:
            _stats(start, n_in, n_out, inp or "stdin", outp or "stdout", mode)
        return 0
    except BrokenPipeError:
        return 0
    except OSError as e:
        print("Error: IO: %s" % e, end="", file=sys.stderr)
        return 1
    except LzfseError as e:
        if e.status == 5:
            print("Error: Buffer overflow", end="", file=sys.stderr)
        elif e.status >= 64:
            print("Error: IO: %s" % e, end="", file=sys.stderr)
        else:
            print("Error: Decode: %s" % e, file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
