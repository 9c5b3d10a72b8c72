"""Developer switches for the tools: every ``--opt name=value`` on the command line becomes a ditto_debug_option call
(DESIGN.md section 9).  Tools only -- the product never reads switches from the command line or the environment."""
import sys


def apply_opts(argv=None):
    from ditto_tts_b200 import _lib
    argv = sys.argv if argv is None else argv
    rest, applied = [], {}
    i = 0
    while i < len(argv):
        if argv[i] == "--lib" and i + 1 < len(argv):      # a differently built library (e.g. the timeline-trace build)
            _lib.LIB_PATH = argv[i + 1]
            i += 2
        elif argv[i] == "--opt" and i + 1 < len(argv):
            k, _, v = argv[i + 1].partition("=")
            _lib.debug_option(k, int(v or 1))
            applied[k] = int(v or 1)
            i += 2
        else:
            rest.append(argv[i])
            i += 1
    argv[:] = rest
    return applied
