// The reference's only automated test (src/cpp/test.cpp:15-60: load a voice, synthesize one sentence,
// require a WAV of at least 10 000 bytes, print OK) restated against the shim.  espeak-ng is a host hook,
// so the sentence arrives pre-phonemized: argv[2] is a JSON-lines file with "phoneme_ids" (and
// "phonemes") like etc/test_sentences/test_en-us.jsonl.
//   usage: test_piper <voice.onnx> <sentences.jsonl> <out.wav> [line]
#include <fstream>
#include <iostream>
#include <sstream>

#include "piper.hpp"

int main(int argc, char* argv[]) {
  if (argc < 4) {
    std::cerr << "usage: test_piper <voice.onnx> <sentences.jsonl> <out.wav> [line]\n";
    return 2;
  }
  const std::string modelPath = argv[1];
  const int wanted = argc > 4 ? std::atoi(argv[4]) : 1;
  try {
    piper::PiperConfig config;
    piper::Voice voice;
    std::optional<piper::SpeakerId> speakerId;
    piper::loadVoice(config, modelPath, modelPath + ".json", voice, speakerId, /*useCuda=*/true);
    piper::initialize(config);

    std::ifstream lines(argv[2]);
    std::string line;
    for (int i = 0; i <= wanted && std::getline(lines, line); ++i) {}
    if (line.empty()) throw std::runtime_error("no such line in the sentences file");
    const piper::json root = minijson::parse(line);

    // 1) ids straight from the fixture  2) ids rebuilt from the fixture's phonemes through phonemes_to_ids
    std::vector<piper::PhonemeId> ids, rebuilt;
    for (const auto& v : root.at("phoneme_ids").arr) ids.push_back(piper::PhonemeId(v.number()));
    if (root.contains("phonemes")) {
      std::vector<piper::Phoneme> ph;
      for (const auto& v : root.at("phonemes").arr) ph.push_back(piper::getCodepoint(v.string()));
      std::map<piper::Phoneme, std::size_t> missing;
      piper::phonemes_to_ids(ph, voice.phonemizeConfig, rebuilt, missing);
      if (rebuilt != ids) throw std::runtime_error("phonemes_to_ids does not reproduce the fixture's phoneme_ids");
    }

    std::stringstream wav;
    piper::SynthesisResult result;
    piper::phonemeIdsToWavFile(voice, ids, wav, result);
    const std::string bytes = wav.str();
    if (bytes.size() < 10000) {
      std::cerr << "ERROR: " << bytes.size() << " bytes of audio, expected >= 10000\n";
      return 1;
    }
    std::ofstream(argv[3], std::ios::binary).write(bytes.data(), std::streamsize(bytes.size()));

    // sentence / silence / callback plumbing through phonemesToAudio
    if (root.contains("phonemes")) {
      std::vector<std::vector<piper::Phoneme>> sentences(2);
      for (const auto& v : root.at("phonemes").arr) sentences[0].push_back(piper::getCodepoint(v.string()));
      sentences[1] = sentences[0];
      std::vector<int16_t> buf;
      int callbacks = 0;
      size_t total = 0;
      piper::SynthesisResult r2;
      piper::phonemesToAudio(config, voice, sentences, buf, r2, [&] { ++callbacks; total += buf.size(); });
      if (callbacks != 2 || !buf.empty() || total == 0) throw std::runtime_error("audio callback contract broken");
      std::cout << "callbacks=" << callbacks << " samples=" << total << " rtf=" << r2.realTimeFactor << "\n";
    }
    piper::terminate(config);
    std::cout << "OK " << bytes.size() << " bytes, infer " << result.inferSeconds << " s, audio "
              << result.audioSeconds << " s, RTF " << result.realTimeFactor << std::endl;
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "ERROR: " << e.what() << std::endl;
    return 1;
  }
}
