#!/bin/bash
OAT_B200_PRE_DEBUG=16 timeout -k 10 120 python tools/tail_probe.py --blobs 60 --frames 64 2>&1 | tail -24
