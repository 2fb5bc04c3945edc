{
  "targets": [{
    "target_name": "aacfb",
    "sources": ["napi/aacfb_napi.c"],
    "include_dirs": ["../../include"],
    "libraries": ["-L<(module_root_dir)/..", "-laacfb", "-Wl,-rpath,<(module_root_dir)/.."]
  }]
}
