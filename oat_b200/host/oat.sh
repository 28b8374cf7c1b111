#!/bin/bash
# oat <component> [TYPE] [IO] [CONFIGURATION] -> exec oat-<component> (oat/libexec/oat:26-40)
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
if [ $# -lt 1 ] || [ "$1" = "help" ] || [ "$1" = "--help" ]; then
    echo "Usage: oat <component> [TYPE] [IO] [CONFIGURATION]"
    echo "Components: framefilt posidet posifilt posicom frameserve posisock clean"
    exit 0
fi
cmd="$1"; shift
if [ ! -x "$here/oat-$cmd" ]; then echo "oat: '$cmd' is not an oat component." >&2; exit 1; fi
exec "$here/oat-$cmd" "$@"
