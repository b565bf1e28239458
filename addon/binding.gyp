{
  "targets": [
    {
      "target_name": "kzgb200",
      "sources": ["kzgb200_napi.cc"],
      "include_dirs": ["../include"],
      "cflags_cc": ["-std=c++17", "-O2"],
      "libraries": ["-L<(module_root_dir)/../kzg_grandsums_study_b200", "-lkzgb200",
                    "-Wl,-rpath,<(module_root_dir)/../kzg_grandsums_study_b200"],
      "defines": ["NAPI_VERSION=8"]
    }
  ]
}
