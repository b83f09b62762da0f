#!/bin/bash
export DDB_TC_ATTN=31
bash profiles/ncu_one.sh told trip_tc_kernel 0 2
