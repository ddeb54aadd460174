// host_cpp_hand_case.cpp — the C++ host layer (include/crgpu.hpp) driving the path end to end.
//
// 1. BarcodeCorrector on the reference's own known-answer cases (lib/rust/barcode/src/corrector.rs:196-277:
//    whitelist {AAAAA, AAGAC, ACGAA, ACGTT}, counts {AAAAA: 100, AAGAC: 11, ACGAA: 2}, Posterior{1.0, 0.95}).
// 2. The three stages on a hand-built GEM well whose count matrix is known by hand (the same case as
//    tests/test_gpu_parity.py::test_umi_chain_and_low_support_hand_case): UMI chain A->B->C with a single hop,
//    a 1:1 tie broken towards the larger UMI (tx_annotation/src/mark_dups.rs:385-391), a UMI seen with two
//    genes (low support), homopolymer and low-quality UMIs dropped (umi/src/info.rs:20-37).
// Exit status 0 and a last line "OK" when every expectation holds.
//
// build: g++ -std=c++17 -Iinclude examples/host_cpp_hand_case.cpp -Lcellranger_b200 -lcrgpu -Wl,-rpath,$PWD/cellranger_b200
#include <cstdio>
#include <string>
#include <vector>

#include "crgpu.hpp"

static int failures = 0;
#define EXPECT(cond)                                                       \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      failures++;                                                          \
    }                                                                      \
  } while (0)

static std::string q(std::initializer_list<int> v) {
  std::string s;
  for (int x : v) s.push_back((char)x);
  return s;
}

int main() {
  try {
    // ---- 1. the corrector plugin seam ----
    {
      crgpu::Posterior strategy;
      strategy.max_expected_barcode_errors = 1.0;
      strategy.bc_confidence_threshold = 0.95;
      crgpu::BarcodeCorrector corrector(crgpu::Whitelist::plain({"AAAAA", "AAGAC", "ACGAA", "ACGTT"}),
                                        {{"AAAAA", 100}, {"AAGAC", 11}, {"ACGAA", 2}}, strategy);
      const auto out = corrector.correct_barcodes({"AAAAT", "ACGAT", "ACGAT", "ACAAA"},
                                                  {q({66, 66, 66, 66, 40}), q({66, 66, 66, 66, 66}),
                                                   q({66, 66, 66, 66, 40}), q({66, 66, 66, 66, 40})});
      EXPECT(out.size() == 4);
      EXPECT(out[0] && out[0]->segment == "AAAAA");  // trivial correction
      EXPECT(!out[1]);                               // pseudo-count kills you
      EXPECT(out[2] && out[2]->segment == "ACGAA");  // quality helps you
      EXPECT(out[3] && out[3]->segment == "AAAAA");  // counts help you
      EXPECT(out[0] && out[0]->state == crgpu::BarcodeSegmentState::ValidAfterCorrection);
    }
    // ---- 2. make_shard -> barcode_correction -> align_and_count ----
    {
      const std::vector<std::string> wl = {"AAAACCCCGGGGTTTT", "ACGTACGTACGTACGT"};
      struct Row {
        std::string bc, umi;
        uint32_t gene;
        int copies;
        char qual;
      };
      const std::vector<Row> rows = {
          {wl[0], "AAAAAAAAAC", 5, 1, 'I'},  // A -> B (count 2) ...
          {wl[0], "AAAAAAAACC", 5, 2, 'I'},  // B -> C (count 3): single hop, B keeps A's read
          {wl[0], "AAAAAAACCC", 5, 3, 'I'},  // C
          {wl[0], "CCCCCCCCCA", 7, 1, 'I'},  // tie 1:1 -> the lexicographically larger CCCCCCCCCG
          {wl[0], "CCCCCCCCCG", 7, 1, 'I'},
          {wl[0], "GATTACAGAT", 3, 4, 'I'},  // the same UMI on two genes: gene 9 is sub-maximal -> low support
          {wl[0], "GATTACAGAT", 9, 1, 'I'},
          {wl[0], "TTTTGGGGCC", 3, 2, 'I'},  // tie across genes -> both low support
          {wl[0], "TTTTGGGGCC", 4, 2, 'I'},
          {wl[0], "GGGGGGGGGG", 3, 5, 'I'},  // homopolymer: invalid UMI
          {wl[0], "ACGTTGCAAC", 3, 2, '*'},  // Q9 < 10: invalid UMI
          {wl[1], "ACGTTGCAAC", 3, 2, 'I'},
      };
      std::vector<uint8_t> r1, q1;
      std::vector<uint32_t> feature;
      for (const auto& r : rows)
        for (int c = 0; c < r.copies; c++) {
          const std::string s = r.bc + r.umi;
          r1.insert(r1.end(), s.begin(), s.end());
          q1.insert(q1.end(), 16, (uint8_t)'I');
          q1.insert(q1.end(), 10, (uint8_t)r.qual);
          feature.push_back(r.gene);
        }
      crgpu::GemWell gw;
      const int w = gw.add_whitelist(crgpu::Whitelist::plain(wl));
      const int lib = gw.add_library(w, crgpu::ChemistryDef::SC3Pv2());
      gw.set_features(std::vector<int32_t>(16, 0));
      gw.add_reads(lib, feature.size(), 26, r1.data(), q1.data(), feature.data());
      gw.run(/*annotate_reads=*/true);
      const crgpu::CountMatrix m = gw.count_matrix();
      EXPECT(m.n_features == 16);
      EXPECT((m.barcodes == std::vector<std::string>{wl[0], wl[1]}));
      EXPECT((m.indptr == std::vector<int64_t>{0, 3, 4}));
      // barcode 0 -> gene 3: GATTACAGAT (1), gene 5: B and C (2), gene 7: CCCCCCCCCG (1); barcode 1 -> gene 3 (1)
      EXPECT((m.indices == std::vector<uint32_t>{3, 5, 7, 3}));
      EXPECT((m.data == std::vector<int32_t>{1, 2, 1, 1}));
      const auto mol = gw.molecules();
      EXPECT(mol.size() == 5);
      uint64_t reads_in_molecules = 0;
      for (const auto& u : mol) reads_in_molecules += u.read_count;
      // GATTACAGAT@3: 4, B: A's read only (its own two moved on to C), C: 3 + 2, CCCCCCCCCG: 1 + 1, barcode 1: 2
      EXPECT(reads_in_molecules == 4 + 1 + 5 + 2 + 2);
      const auto summary = gw.barcode_summary(lib);
      EXPECT(summary.size() == 2);
      if (summary.size() == 2) {
        EXPECT(summary[0].reads == 24 && summary[1].reads == 2);
        EXPECT(summary[0].umis == 4 && summary[1].umis == 1);
        EXPECT(summary[0].umi_corrected_reads == 4);               // A's read, B's two reads, CCCCCCCCCA's read
        EXPECT(summary[0].candidate_dup_reads == 17 - 1 - 4);      // 17 reads with DupInfo, 5 of them low support
      }
      const auto reads = gw.reads(0);
      EXPECT(reads.bc_state.size() == feature.size());
      EXPECT(reads.bc_state[0] == crgpu::BarcodeSegmentState::ValidBeforeCorrection);
    }
  } catch (const crgpu::Error& e) {
    std::printf("crgpu::Error: %s\n", e.what());
    return 2;
  }
  if (failures) return 1;
  std::printf("OK\n");
  return 0;
}
