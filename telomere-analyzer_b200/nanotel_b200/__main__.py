"""python -m nanotel_b200 -i reads.fastq.gz --save_path out --patterns "TTAGGG"   (NanoTel.R's command line)"""
from .nanotel import main

raise SystemExit(main())
