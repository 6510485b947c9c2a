#!/bin/bash
python tools/e2e_probe.py 10000
