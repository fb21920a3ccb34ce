from kindel_b200.cli import main

raise SystemExit(main())
