#!/bin/bash
# rebuild libcapdec.so regardless of the caller's working directory
cd "$(dirname "$0")/.." && python -c "from myimagecaptioningmodel_b200 import _build; print(_build.build())"
