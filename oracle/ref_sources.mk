# Reference translation units that make up the population-genotype hot path.
# They are compiled IN PLACE from $(REF) (= /root/reference); nothing is copied.
# List follows SURVEY.md section 8c (link recipe).
REF_TUS := \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_calc.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_freq.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_locus.cpp \
  kga_analytic/kga_PfEMP/kga_analysis_PfEMP_FWS.cpp \
  kga_analytic/kga_PfEMP/kga_analysis_PfEMP_heterozygous.cpp \
  $(patsubst $(REF)/%,%,$(wildcard $(REF)/kgl_genomics/kgl_variant_db/*.cpp)) \
  $(patsubst $(REF)/%,%,$(wildcard $(REF)/kgl_genomics/kgl_variant_filter/*.cpp)) \
  $(patsubst $(REF)/%,%,$(wildcard $(REF)/kgl_genomics/kgl_evidence/*.cpp)) \
  $(patsubst $(REF)/%,%,$(wildcard $(REF)/kgl_genomics/kgl_sequence/*.cpp)) \
  kel_app/kel_exec_env.cpp kel_app/kel_logging.cpp kel_app/kel_logging_stream.cpp \
  kel_utility/kel_utility.cpp kel_utility/kel_mem_alloc.cpp kel_utility/kel_search.cpp kel_utility/kel_interval_set.cpp \
  kgl_genomics/kgl_parser/kgl_variant_factory_vcf_parse_info.cpp kgl_genomics/kgl_parser/kgl_data_file_type.cpp \
  kgl_app/kgl_runtime.cpp kgl_app/kgl_runtime_resource.cpp

# Additional reference TUs for the plugin harness (the whole INBREED analysis, its PED resource and CSV writer).
PLUGIN_TUS := \
  kga_analytic/kga_inbreed/kga_analysis_inbreed.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_diploid.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_execute.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_output.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_args.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_synthetic.cpp \
  kga_analytic/kga_inbreed/kga_analysis_inbreed_syngen.cpp \
  kgl_genomics/kgl_parser/kgl_hsgenealogy_parser.cpp \
  kgl_genomics/kgl_parser/kgl_hsgenome_aux.cpp \
  kgl_genomics/kgl_parser/kgl_square_parser.cpp \
  kgl_genomics/kgl_parser/kgl_pf7_sample_parser.cpp kgl_genomics/kgl_parser/kgl_pf7_fws_parser.cpp \
  kgl_genomics/kgl_parser/kgl_variant_vcf_impl.cpp kgl_genomics/kgl_parser/kgl_variant_factory_readvcf_impl.cpp \
  kgl_genomics/kgl_parser/kgl_variant_factory_1000_impl.cpp \
  kel_io/kel_mt_buffer.cpp kel_io/kel_basic_io.cpp

REF_INCLUDES := contrib/edlib kel_utility kel_thread kel_io kgl_genomics kel_app kel_math kgl_app \
  kgl_genomics/kgl_parser kgl_genomics/kgl_evidence kgl_genomics/kgl_sequence kgl_genomics/kgl_database \
  kgl_genomics/kgl_classification kgl_genomics/kgl_genome kgl_genomics/kgl_genome_io kgl_genomics/kgl_variant_db \
  kgl_genomics/kgl_variant_filter kgl_genomics/kgl_variant_analysis kgl_genomics/kgl_literature \
  kgl_genomics/kgl_mutation kgl_genomics/kgl_legacy kol_ontology kol_ontology/kgl_ontology \
  kga_analytic/kga_inbreed kga_analytic/kga_sequence_analysis kga_analytic/kga_info kga_analytic/kga_PfEMP \
  kga_analytic/kga_analysis_library kga_analytic/kga_literature kga_analytic/kga_mutation \
  kga_analytic/kga_template_analysis kga_analytic contrib/rapidjson/include contrib/rapidxml
