#!/usr/bin/env bash
# Renders the cross-check cases with the real reference and stores its outputs as golden fixtures:
#   tools/dotnet_crosscheck/run.sh            (needs the .NET 9 SDK; GRAPHAUDIO_CORE=<path to GraphAudio.Core.csproj> overrides the reference)
# Afterwards `python -m pytest tests/test_reference_crosscheck.py` compares the CPU oracle (and, with -m gpu, the CUDA path) with them.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
if ! command -v dotnet >/dev/null 2>&1; then
  echo "dotnet not found: the reference (C#/.NET 9) cannot be built here" >&2
  exit 3
fi
python "$here/dump_cases.py" "$here/cases"
core="${GRAPHAUDIO_CORE:-/root/reference/GraphAudio.Core/GraphAudio.Core.csproj}"
dotnet run -c Release --project "$here/Crosscheck.csproj" -p:GraphAudioCore="$core" -- "$here/cases" "$root/tests/golden"
ls -l "$root"/tests/golden/ref_*.f32
