#!/bin/bash
exec bash "$(dirname "$0")/session_multi.sh" 8
