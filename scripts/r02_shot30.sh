#!/bin/bash
bash scripts/r02_shot28.sh
bash scripts/r02_shot27.sh 2>&1 | tail -60
