#!/usr/bin/env bash
# Installs the JLL override for OSQP.jl.  Usage: install.sh [--depot DIR] [--project DIR]
#   --depot DIR    Julia depot to receive artifacts/Overrides.toml   (default: ${JULIA_DEPOT_PATH%%:*} or ~/.julia)
#   --project DIR  Julia project to receive LocalPreferences.toml     (default: none -> mechanism A only)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
prefix="$here/prefix"
depot="${JULIA_DEPOT_PATH:-$HOME/.julia}"; depot="${depot%%:*}"
project=""
while [ $# -gt 0 ]; do
  case "$1" in
    --depot) depot="$2"; shift 2 ;;
    --project) project="$2"; shift 2 ;;
    *) echo "unknown argument $1" >&2; exit 2 ;;
  esac
done
[ -e "$prefix/lib/libosqp.so" ] || { echo "build the engine first (python -c 'import __graft_entry__ as g; g.build()')" >&2; exit 1; }
mkdir -p "$depot/artifacts"
if [ -f "$depot/artifacts/Overrides.toml" ] && ! grep -q 9c4f68bf-6205-5545-a508-2878b064d984 "$depot/artifacts/Overrides.toml"; then
  sed "s|@PREFIX@|$prefix|; /^#/d" "$here/artifacts/Overrides.toml" >> "$depot/artifacts/Overrides.toml"
else
  sed "s|@PREFIX@|$prefix|" "$here/artifacts/Overrides.toml" > "$depot/artifacts/Overrides.toml"
fi
echo "wrote $depot/artifacts/Overrides.toml"
if [ -n "$project" ]; then
  sed "s|@PREFIX@|$prefix|" "$here/LocalPreferences.toml" > "$project/LocalPreferences.toml"
  echo "wrote $project/LocalPreferences.toml"
fi
