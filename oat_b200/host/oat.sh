#!/bin/bash
# oat <component> [TYPE] [IO] [CONFIGURATION] -> exec oat-<component> (oat/libexec/oat:26-40)
# oat mps start|stop: a graph is one OS process per component; without the CUDA MPS daemon the GPU time-slices between
# their contexts (a 3-component chain of device frames: 1.4 k frames/s; under MPS: 15 k, profiles/r02t_graph_bench_mps.txt).
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
if [ $# -lt 1 ] || [ "$1" = "help" ] || [ "$1" = "--help" ]; then
    echo "Usage: oat <component> [TYPE] [IO] [CONFIGURATION]"
    echo "   or: oat mps start|stop      (share the GPU between the components of a graph: start before the graph)"
    echo "Components: framefilt posidet posifilt posicom frameserve buffer posisock clean"
    exit 0
fi
cmd="$1"; shift
if [ "$cmd" = "mps" ]; then
    case "$1" in
        start) exec nvidia-cuda-mps-control -d ;;
        stop) echo quit | nvidia-cuda-mps-control; exit $? ;;
        *) echo "oat mps: start or stop" >&2; exit 1 ;;
    esac
fi
if [ ! -x "$here/oat-$cmd" ]; then echo "oat: '$cmd' is not an oat component." >&2; exit 1; fi
exec "$here/oat-$cmd" "$@"
