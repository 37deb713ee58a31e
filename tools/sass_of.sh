#!/bin/bash
# usage: tools/sass_of.sh <regex on mangled function name> -> clean SASS listing on stdout
cuobjdump -sass audiolazy_b200/_native/libalz_b200.so | awk -v pat="$1" '/Function : /{f=($0 ~ pat)} f{print}' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]{4})\*\/\s+/\1 /; s/\s*\/\*.*$//'
