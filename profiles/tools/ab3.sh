timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
AB_ENVS="LUZRT_TRI_VOTE_DIV=0 LUZRT_TRI_VOTE_DIV=1 LUZRT_TRI_VOTE_DIV=2 LUZRT_TRI_VOTE_DIV=3 LUZRT_TRI_VOTE_DIV=4 LUZRT_TRI_VOTE_DIV=8" bash profiles/ab.sh ab3 build/variants/libluzrt_vote.so
AB_CONFIGS="c3 c2" AB_ENVS="LUZRT_TRI_VOTE_DIV=0 LUZRT_TRI_VOTE_DIV=3" bash profiles/ab.sh ab3b build/variants/libluzrt_split.so build/variants/libluzrt_vote_slowgen.so
